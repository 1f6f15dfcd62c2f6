#!/bin/bash
# grouped weight stages for the 96-channel tap-reuse layers (A/B), FLAME v3 timing, conv-net parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flame.py tests/test_gpu_net.py -q -x > gpurun_out/r2b_wgroup_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_wgroup_tests.log
timeout 300 python tools/bench_flame.py r2b_dmma3 2>&1 | tail -7
for wg in 0 1; do
  VGGHEADS_B200_WGROUP=$wg timeout 600 python tools/profile_ops.py 64 640 r2b_wg$wg > gpurun_out/r2b_ops_wg$wg.log 2>&1
  head -1 gpurun_out/ops_r2b_wg$wg.txt; grep -E "stage1.csp.b|flame_decode" gpurun_out/ops_r2b_wg$wg.txt | cut -c1-140
done
for wg in 0 1; do
  VGGHEADS_B200_WGROUP=$wg timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_wg$wg.json 2> gpurun_out/r2b_bench_wg$wg.err; echo "bench wg$wg rc=$?"
done
python - <<'PY'
import json
for wg in (0, 1):
    d = json.loads(open(f"gpurun_out/r2b_bench_wg{wg}.json").read().strip().splitlines()[-1])
    print("wgroup", wg, round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"], 3), "frac_step", round(d["roofline"]["frac_step"], 3), d["clocks"]["sm_mhz"])
PY

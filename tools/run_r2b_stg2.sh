#!/bin/bash
# residual prefetch with two staging tiles (96-channel tap-reuse layers): parity tests + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -q -x > gpurun_out/r2b_stg2_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_stg2_tests.log
for v in 0 1; do
  VGGHEADS_B200_STG2=$v timeout 600 python tools/profile_ops.py 64 640 r2b_stg$v > gpurun_out/r2b_ops_stg$v.log 2>&1
  head -1 gpurun_out/ops_r2b_stg$v.txt; grep -E "stage1.csp.b" gpurun_out/ops_r2b_stg$v.txt | cut -c1-140
done
for v in 0 1; do
  VGGHEADS_B200_STG2=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_stg$v.json 2> gpurun_out/r2b_bench_stg$v.err; echo "bench stg$v rc=$?"
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads(open(f"gpurun_out/r2b_bench_stg{v}.json").read().strip().splitlines()[-1])
    print("stg2", v, round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"], 3), "frac_step", round(d["roofline"]["frac_step"], 3), d["clocks"]["sm_mhz"])
PY

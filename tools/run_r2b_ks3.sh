#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -q -x -k "autotuned or variants or every_buffer" > gpurun_out/r2b_ks3_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2b_ks3_tests.log
SECONDS=0; timeout 600 python tools/profile_ops.py 64 640 r2b_ks3 > gpurun_out/r2b_ops_ks3.log 2>&1; echo "profile wall ${SECONDS}s"
head -1 gpurun_out/ops_r2b_ks3.txt; grep -E "bk32" gpurun_out/ops_r2b_ks3.txt | cut -c1-140
SECONDS=0; timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_ks3.json 2> gpurun_out/r2b_bench_ks3.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_ks3.json").read().strip().splitlines()[-1])
print(round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"], 3), "frac_step", round(d["roofline"]["frac_step"], 3), d["clocks"]["sm_mhz"])
PY

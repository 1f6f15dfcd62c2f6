#!/bin/bash
# 2-GPU sanity check of the final code: two-rank gather tests + the driver's scaling command at N=2 (default transport) and N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -x -q > gpurun_out/r2b_parallel_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2b_parallel_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -c 400 gpurun_out/r2b_bench_2gpu.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_1gpu_same_box.json 2> gpurun_out/r2b_bench_1gpu_same_box.err; echo "bench 1gpu rc=$?"
python - <<'PY'
import json
for f in ("r2b_bench_1gpu_same_box","r2b_bench_2gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "img/s  e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), d.get("gather"), d["clocks"]["sm_mhz"])
    except Exception as ex:
        print(f, "no line", ex)
PY

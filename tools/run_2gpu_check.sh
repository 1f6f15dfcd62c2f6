#!/bin/bash
# 2-GPU check of the multi-GPU path: gpu tests of the gather, then the driver's scaling command at N=2 (both transports) and N=1.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parallel.py -x -q > gpurun_out/r2_parallel_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_parallel_tests.log
tail -5 gpurun_out/r2_parallel_tests.log
for g in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --gather $g > gpurun_out/r2_bench_2gpu_$g.json 2> gpurun_out/r2_bench_2gpu_$g.err; echo "bench 2gpu $g rc=$?"
  tail -c 600 gpurun_out/r2_bench_2gpu_$g.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench 1gpu rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_1gpu","r2_bench_2gpu_peer","r2_bench_2gpu_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "img/s  e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), d.get("gather"))
    except Exception as ex:
        print(f, "no line", ex)
PY

#!/bin/bash
# 8-GPU box: the driver's scaling commands at N = 8, 4, 2, 1 back to back (+ the two-rank gather tests).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; echo "N=$n rc=$?"
  grep -E "^rank|Error|error" gpurun_out/r2_scale_n$n.err | head -5
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 --gather nccl > gpurun_out/r2_scale_n8_nccl.json 2> gpurun_out/r2_scale_n8_nccl.err; echo "N=8 nccl rc=$?"
timeout 900 python -m pytest tests/test_gpu_parallel.py -q -k "two_rank or two_ranks" > gpurun_out/r2_parallel_tests_8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_parallel_tests_8.log
python - <<'PY'
import json
base=None
for n in (1,2,4,8,"8_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r2_scale_n{n}.json").read().strip().splitlines()[-1])
        if n==1: base=d["value"]
        k=int(str(n).split("_")[0])
        print(f"N={n}: {d['value']:.0f} img/s (eff {d['value']/(k*base):.3f}) e2e {d['e2e']['value']:.0f} ms/step {d['ms_per_step']:.3f} {d.get('gather',{}).get('transport','')[:40]}")
    except Exception as ex:
        print(n, "no line", ex)
PY

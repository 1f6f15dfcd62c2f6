#!/bin/bash
# run-time parameter sweep on one GPU: per-GPU batch x detector handles in flight
mkdir -p gpurun_out
for cfg in "32 2" "64 2" "64 3" "128 2" "64 1"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --per-gpu-batch $1 --engines $2 --no-cpu-baseline --no-extras > gpurun_out/r2_sweep_b$1_e$2.json 2> gpurun_out/r2_sweep_b$1_e$2.err
  python - "$1" "$2" <<'PY'
import json,sys
b,e=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2_sweep_b{b}_e{e}.json").read().strip().splitlines()[-1])
    print(f"B={b} engines={e}: {d['value']:.0f} img/s e2e {d['e2e']['value']:.0f} frac_serial {d['roofline']['frac_serial']:.3f} frac_step {d['roofline']['frac_step']:.3f} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as ex:
    print(b,e,"failed",ex); print(open(f"gpurun_out/r2_sweep_b{b}_e{e}.err").read()[-800:])
PY
done

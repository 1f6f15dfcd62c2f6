#!/bin/bash
# what the driver runs at round end, on one GPU: gpu tests, smoke(), the reference arm (short) and bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_final_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; echo "reference arm rc=$?"
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_bench_1gpu.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_final_reference_arm.json").read().strip().splitlines()[-1])
print(round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"],3), "frac_step", round(d["roofline"]["frac_step"],3), d["clocks"]["sm_mhz"], "dense", round(d["dense_heads"].get("value",0)), "cfg4", round(d["config4_shard"].get("value",0)), "cpu", round(d["cpu_baseline"]["value"],1), "ref arm", round(r["value"],1))
print({k:v for k,v in d["parity"].items() if k!="end_to_end" and k!="checker"})
PY

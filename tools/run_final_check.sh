#!/bin/bash
# what the driver runs at round end, on one GPU: gpu tests, smoke(), the reference arm (short) and bench.py; plus the evidence that
# goes with the final code: ncu launch list of one tuned batch-64 step and the per-op table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/r2b_final_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_final_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2b_final_reference_arm.json 2> gpurun_out/r2b_final_reference_arm.err; echo "reference arm rc=$?"
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_final_bench_1gpu.json 2> gpurun_out/r2b_final_bench_1gpu.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_final_bench_1gpu.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2b_final_reference_arm.json").read().strip().splitlines()[-1])
print(round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), d["dtype"], "frac_serial", round(d["roofline"]["frac_serial"],3), "frac_step", round(d["roofline"]["frac_step"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "dense", round(d["dense_heads"].get("value",0)), "cfg4", round(d["config4_shard"].get("value",0)), "cpu", round(d["cpu_baseline"]["value"],1), "ref arm", round(r["value"],1))
print({k:v for k,v in d["parity"].items() if k!="end_to_end" and k!="checker"})
e=d["parity"]["end_to_end"]
for k in e:
    if isinstance(e[k], dict) and "boxes_max_abs_err_px" in e[k]:
        print("  ", k, {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in e[k].items() if kk != "stage_rel_err"})
PY
timeout 300 python tools/bench_flame.py r2b_final 2>&1 | tail -7
timeout 600 python tools/profile_ops.py 64 640 r2b_final > gpurun_out/r2b_final_ops.log 2>&1; head -2 gpurun_out/ops_r2b_final.txt
VGGHEADS_B200_SPARSE_HEADS=1 timeout 900 ncu --profile-from-start off --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
  --log-file gpurun_out/r2b_final_ncu_launches.csv python tools/ncu_target.py 64 tuned > gpurun_out/r2b_final_ncu_target.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2b_final_ncu_launches.csv

#!/bin/bash
# 1-GPU check: full gpu test-suite (single-GPU part), per-op profile, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parallel.py::test_two_rank_gather --deselect tests/test_gpu_parallel.py::test_bench_run_ours_two_ranks > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|FAILED|heads vs reference|decoded:|per-stage|parity=" gpurun_out/r2_gputests.log | cut -c1-700 | tail -40
timeout 600 python tools/profile_ops.py 64 640 r2_b64 > gpurun_out/r2_ops_b64.log 2>&1; head -12 gpurun_out/ops_r2_b64.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_c.json 2> gpurun_out/r2_bench_1gpu_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu_c.json").read().strip().splitlines()[-1])
print(round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"],3), "frac_step", round(d["roofline"]["frac_step"],3), d["clocks"]["sm_mhz"])
PY
if [ -n "$NCU" ]; then
  VGGHEADS_B200_SPARSE_HEADS=1 timeout 900 ncu --profile-from-start off --clock-control none --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --log-file gpurun_out/r2_ncu_launches.csv python tools/ncu_target.py 64 tuned > gpurun_out/r2_ncu_target.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/r2_ncu_launches.csv
fi

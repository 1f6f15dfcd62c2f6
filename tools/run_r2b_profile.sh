#!/bin/bash
# round 2, second session: FP64 rate probe, per-op table, ncu launch list of one tuned batch-64 step (fp16 storage, CTA pairs),
# one --set full capture of the first 96-channel tap-reuse launch (stage1.csp.b0.cv1)
mkdir -p gpurun_out
./tools/microbench/fp64_rate > gpurun_out/r2b_fp64_rate.txt 2>&1; cat gpurun_out/r2b_fp64_rate.txt
timeout 600 python tools/profile_ops.py 64 640 r2b_b64 > gpurun_out/r2b_ops_b64.log 2>&1; head -3 gpurun_out/ops_r2b_b64.txt
VGGHEADS_B200_SPARSE_HEADS=1 timeout 900 ncu --profile-from-start off --clock-control none --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
  --log-file gpurun_out/r2b_ncu_launches.csv python tools/ncu_target.py 64 tuned > gpurun_out/r2b_ncu_target.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2b_ncu_launches.csv
VGGHEADS_B200_SPARSE_HEADS=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k 'regex:conv_igemm_swap_kernel<32, true' -c 1 -o gpurun_out/r2b_ncu_full_stage1_xr32 -f python tools/ncu_target.py 64 untuned > gpurun_out/r2b_ncu_full_stage1.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null; tail -3 gpurun_out/r2b_ncu_full_stage1.log

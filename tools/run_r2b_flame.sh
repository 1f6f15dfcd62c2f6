#!/bin/bash
# FLAME decode on the FP64 tensor pipe: parity tests, stand-alone timing, and one --set full capture each of the FLAME kernel
# and of the first 96-channel tap-reuse conv launch (stage1.csp.b0.cv1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flame.py tests/test_gpu_net.py tests/test_gpu_mesh.py -q -x > gpurun_out/r2b_flame_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_flame_tests.log
timeout 300 python tools/bench_flame.py r2b_dmma 2>&1 | tail -8
VGGHEADS_B200_SPARSE_HEADS=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k 'regex:flame_decode_kernel' -c 1 -o gpurun_out/r2b_ncu_full_flame_dmma -f python tools/ncu_target.py 64 untuned > gpurun_out/r2b_ncu_full_flame.log 2>&1; echo "ncu flame rc=$?"
VGGHEADS_B200_SPARSE_HEADS=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k 'regex:conv_igemm_swap_kernel<\(int\)32, \(bool\)1' -c 1 -o gpurun_out/r2b_ncu_full_stage1_xr32 -f python tools/ncu_target.py 64 untuned > gpurun_out/r2b_ncu_full_stage1.log 2>&1; echo "ncu stage1 rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null

#!/bin/bash
# usage: tools/ncu_full_kernel.sh <kernel regex> <tag> [batch]   - one --set full capture (with source) of one kernel of a tuned step
mkdir -p gpurun_out
VGGHEADS_B200_SPARSE_HEADS=1 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$1 -c 1 \
  -o gpurun_out/r2_ncu_full_$2 -f python tools/ncu_target.py ${3:-64} untuned > gpurun_out/r2_ncu_full_$2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_ncu_full_$2.ncu-rep

#!/bin/bash
# ncu --set full of the non-GEMM, non-attention kernels of one training step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"ln_|colsum|rowdot|transpose_sum|patchify|latw|cast_" -c 14 -f -o gpurun_out/r1_misc \
  python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/ncu_misc.log; ls -la gpurun_out

#!/bin/bash
# ncu --set full of the attention-backward, LN-backward and column-sum kernels of one training step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"${KERNELS:-attn_tc_bwd2|attn_rowdot|ln_bwd_rowwarp|colsum}" -c ${COUNT:-8} -f -o gpurun_out/r1_b_bwd_kernels \
  python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/ncu_misc.log; ls -la gpurun_out/*.ncu-rep

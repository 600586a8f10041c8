#!/bin/bash
# ncu --set full captures for profiles/: plain + GELU GEMM, attention forward / backward (final round-1 kernels)
mkdir -p gpurun_out
for c in fc2 gelu wgrad; do
  timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r1_final_gemm_$c python tools/gemm_one.py $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c exit $?"
done
timeout 600 ncu --set full --clock-control none -k regex:"attn_tc_|attn_rowdot" -s 3 -c 3 -f -o gpurun_out/r1_final_attn python tools/attn_phases.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep

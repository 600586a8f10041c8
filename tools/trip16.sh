#!/bin/bash
# ncu --set full captures (with source counters) of the GELU / DGELU GEMMs and the attention backward
mkdir -p gpurun_out
for c in gelu dgelu; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r1_gemm_$c python tools/gemm_one.py $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c exit $?"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc_bwd_kernel -s 1 -c 1 -f -o gpurun_out/r1_attn_bwd python tools/attn_phases.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep

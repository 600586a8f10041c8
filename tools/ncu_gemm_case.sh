#!/bin/bash
mkdir -p gpurun_out
for c in ${CASES:-gelu}; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r1_gemm_${c}_v5 python tools/gemm_one.py $c > gpurun_out/ncu_$c.log 2>&1; echo "ncu $c exit $?"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/prof_gelu python bench.py --profile-step > gpurun_out/ncu_b.log 2>&1; echo "ncu gelu exit $?"; tail -3 gpurun_out/ncu_b.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# DRAM bytes of every tcgen05 GEMM launch of one training step (for roofline.traffic)
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
  -k regex:gemm_tc_kernel --csv --log-file gpurun_out/r1_b_gemm_traffic.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r1_b_gemm_traffic.csv

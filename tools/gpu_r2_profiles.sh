#!/bin/bash
# round-2 evidence for profiles/: launch list + GEMM DRAM traffic of one profiled step, ncu --set full of the new attention kernels,
# the attention sweep (BASELINE config 5), the config-4 workload, the reference arm
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --profile-step --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"; wc -l gpurun_out/r2_launches.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
  -k regex:gemm_tc_kernel --csv --log-file gpurun_out/r2_gemm_traffic.csv python bench.py --profile-step --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_fwd3|attn_tc_bwd3|attn_rowdot3" -s 3 -c 3 -f -o gpurun_out/r2_attn python tools/attn_phases.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_gemm_qknorm python tools/gemm_one.py qknorm > gpurun_out/ncu_qknorm.log 2>&1; echo "ncu qknorm exit $?"
timeout 900 python bench.py --sweep attn --steps 5 --warmup 2 > gpurun_out/r2_attn_sweep.jsonl 2> gpurun_out/r2_attn_sweep.err; echo "sweep exit $?"; wc -l gpurun_out/r2_attn_sweep.jsonl
timeout 900 python bench.py --workload config4 --batch-per-gpu 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_config4_b2.json 2> gpurun_out/r2_config4_b2.err; echo "config4 exit $?"; cut -c1-600 gpurun_out/r2_config4_b2.json
timeout 900 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/r2_bench_ref.json
timeout 900 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/r2_bench_ref2.json 2> gpurun_out/r2_bench_ref2.err; cut -c1-200 gpurun_out/r2_bench_ref2.json

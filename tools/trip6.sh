#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | cut -c1-600 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile-step > gpurun_out/ncu_a.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 8 -c 6 -o gpurun_out/prof_gemm_r1 python bench.py --profile-step > gpurun_out/ncu_b.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc -s 2 -c 2 -o gpurun_out/prof_attn_r1 python bench.py --profile-step > gpurun_out/ncu_c.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | cut -c1-300 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1b.csv python bench.py --profile-step > gpurun_out/ncu_a.log 2>&1; echo "ncu list exit $?"

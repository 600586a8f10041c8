#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | cut -c1-300 | tail -8
timeout 300 python tools/gemm_bench.py 2>&1 | tail -8

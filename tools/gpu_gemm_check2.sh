#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "gemm tests exit $?"; tail -3 gpurun_out/t_gemm.log
timeout 300 python tools/gemm_bench.py 2>&1 | tail -11

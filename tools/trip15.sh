#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "tcgen05" --timeout 200 -x > gpurun_out/t_pair.log 2>&1; echo "pair tests exit $?"
grep -E "^E  |^FAILED|passed|failed|timed out|swinb200:" gpurun_out/t_pair.log | cut -c1-300 | tail -12
echo "== pair on"; timeout 200 python tools/gemm_bench.py 2>&1 | tail -6
echo "== pair off"; SWINB200_GEMM_PAIR=0 timeout 200 python tools/gemm_bench.py 2>&1 | tail -6

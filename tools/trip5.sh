#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | cut -c1-600 | tail -30
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r1_b.err; cat gpurun_out/bench_r1_b.json

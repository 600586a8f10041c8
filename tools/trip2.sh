#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/t_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json

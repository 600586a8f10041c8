#!/bin/bash
# round-2 check: GPU tests (all), attention phase breakdown, bench (with eager + cpu baselines)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -25 gpurun_out/t_gpu.log
timeout 300 python tools/attn_phases.py > gpurun_out/attn_phases.log 2>&1; cat gpurun_out/attn_phases.log
timeout 900 python bench.py > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err; echo "bench exit $?"; cat gpurun_out/r2_a_bench.json | cut -c1-6000; tail -3 gpurun_out/r2_a_bench.err

#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -2 gpurun_out/t_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r1_b_bench.json 2> gpurun_out/r1_b_bench.err; echo "bench exit $?"; cat gpurun_out/r1_b_bench.json | cut -c1-1500
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_b_bench_ref.json 2> gpurun_out/r1_b_bench_ref.err; echo "ref exit $?"; cat gpurun_out/r1_b_bench_ref.json | cut -c1-800

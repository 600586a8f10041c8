#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -2 gpurun_out/t_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_b_bench.json 2> gpurun_out/r2_b_bench.err; echo "bench exit $?"; cut -c1-700 gpurun_out/r2_b_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_b_bench_ref.json 2> gpurun_out/r2_b_bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/r2_b_bench_ref.json

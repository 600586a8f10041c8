#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "tests exit $?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/t_all.log | cut -c1-600 | tail -12
timeout 300 python tools/attn_phases.py 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r1_c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_families'], d['roofline']['achieved'])"

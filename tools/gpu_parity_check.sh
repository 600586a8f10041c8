#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo "pytest attention exit $?"; tail -3 gpurun_out/t_attn.log
timeout 1500 python -m pytest tests/test_parity_headline_gpu.py tests/test_widening_gpu.py tests/test_model_gpu.py -q -s > gpurun_out/t_parity.log 2>&1; echo "pytest parity exit $?"; grep -E "passed|failed|FAILED|floor|^E " gpurun_out/t_parity.log | cut -c1-400 | tail -30
timeout 300 python tools/bwd3_time.py 2>&1 | tail -2

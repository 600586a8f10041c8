#!/bin/bash
# attention kernels: parity tests + phase breakdown / timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo "pytest attention exit $?"; tail -15 gpurun_out/t_attn.log
timeout 300 python tools/attn_phases.py > gpurun_out/attn_phases.log 2>&1; cat gpurun_out/attn_phases.log

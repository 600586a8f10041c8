#!/bin/bash
# attention: parity tests, phases, and one ncu --set full capture (with source) of the backward kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > gpurun_out/t_attn.log 2>&1; echo "pytest attention exit $?"; tail -5 gpurun_out/t_attn.log
timeout 300 python tools/attn_phases.py > gpurun_out/attn_phases.log 2>&1; cat gpurun_out/attn_phases.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_bwd3" -s 1 -c 1 -f -o gpurun_out/r2_attn_bwd3 python tools/attn_phases.py > gpurun_out/ncu_bwd3.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep | tail -2

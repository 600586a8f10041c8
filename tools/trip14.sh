#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ln_residual_bwd" -s 1 -c 1 -o gpurun_out/prof_lnb python bench.py --profile-step > gpurun_out/ncu_b.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_b.log

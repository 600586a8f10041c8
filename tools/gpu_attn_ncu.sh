#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_bwd3" -s 1 -c 1 -f -o gpurun_out/r2_attn_bwd3 python tools/attn_phases.py > gpurun_out/ncu_bwd3.log 2>&1; echo "ncu exit $?"

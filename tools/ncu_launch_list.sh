#!/bin/bash
# launch list of one training step (per-kernel durations)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r1_b_launches.csv python bench.py --profile-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
wc -l gpurun_out/r1_b_launches.csv

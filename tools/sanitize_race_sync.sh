#!/bin/bash
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/$tool.log \
    python -m pytest tests/test_kernels_gpu.py -x -q --timeout 1100 > gpurun_out/${tool}_pytest.log 2>&1
  echo "$tool exit $?"; tail -1 gpurun_out/${tool}_pytest.log; grep -c "hazard\|Barrier error\|Error" gpurun_out/$tool.log; tail -3 gpurun_out/$tool.log
done

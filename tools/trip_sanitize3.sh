#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck_ln.log \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "ln_residual" --timeout 800 > gpurun_out/racecheck_ln_pytest.log 2>&1
echo "racecheck exit $?"; tail -1 gpurun_out/racecheck_ln_pytest.log; grep "Error\|SUMMARY" gpurun_out/racecheck_ln.log | cut -c1-200 | head

#!/bin/bash
# compute-sanitizer memcheck over the kernel parity tests (small shapes)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "${1:-not full}" --timeout 1400 > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck exit $?"
tail -3 gpurun_out/memcheck_pytest.log
grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log

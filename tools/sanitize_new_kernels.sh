#!/bin/bash
# racecheck + synccheck over the round-2 kernels (tiled attention, fused weight + bias gradient, fused LN epilogue) at small shapes
mkdir -p gpurun_out
SEL='geometry_sweep or (linear_wgrad and 1000) or (linear_wgrad and 5000) or (linear_ln_residual and 300) or window_attention_tcgen05'
for tool in racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r2_$tool.log \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$SEL" --timeout 1400 > gpurun_out/r2_${tool}_pytest.log 2>&1
  echo "$tool exit $?"; tail -1 gpurun_out/r2_${tool}_pytest.log; grep -c "hazard\|Barrier error\|Error" gpurun_out/r2_$tool.log; tail -3 gpurun_out/r2_$tool.log
done

#!/bin/bash
# first GPU trip: kernel parity (CUDA-core paths), tcgen05 GEMM probe, model parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "not tcgen05" -x --timeout 600 > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; echo "probe exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "tcgen05" --timeout 300 > gpurun_out/t_tc.log 2>&1; echo "tcgen05 tests exit $?" | tee -a gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_model_gpu.py -q -k "not bf16-" --timeout 900 > gpurun_out/t_model.log 2>&1; echo "model exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_model_gpu.py -q -k "bf16-" --timeout 300 > gpurun_out/t_model_tc.log 2>&1; echo "model tc exit $?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_kernels.log; cat gpurun_out/gemm_probe.log | tail -60; tail -5 gpurun_out/t_tc.log; tail -5 gpurun_out/t_model.log; tail -5 gpurun_out/t_model_tc.log

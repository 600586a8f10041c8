#!/bin/bash
# multi-GPU: NCCL DDP gradient-equivalence test + bench at N = 1..$1 (both arms at N>1 would repeat the CPU run: ours only)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_parity_headline_gpu.py -q -x -k "ddp" -s > gpurun_out/t_ddp.log 2>&1; echo "ddp tests exit $?"; grep -E "passed|failed|skipped|ddp_" gpurun_out/t_ddp.log | tail -4
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
    else
      NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline > gpurun_out/r2_scale_n$n.out 2> gpurun_out/r2_scale_n$n.err
      grep '^{"metric' gpurun_out/r2_scale_n$n.out > gpurun_out/r2_scale_n$n.json
      grep -c "NCCL INFO" gpurun_out/r2_scale_n$n.out gpurun_out/r2_scale_n$n.err | tr '\n' ' '; grep -m2 -h "nranks\|NVLS" gpurun_out/r2_scale_n$n.out gpurun_out/r2_scale_n$n.err | cut -c1-200
    fi
    echo "N=$n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_scale_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
  fi
done

#!/bin/bash
# scaling run on one box with $1 GPUs: N = 1, 2, 4, 8 (as many as fit)
mkdir -p gpurun_out
NG=$1
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_r2_n$n.json 2> gpurun_out/scale_r2_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_r2_n$n.json 2> gpurun_out/scale_r2_n$n.err
    fi
    echo "N=$n exit $?"
    python -c "
import json
txt=open('gpurun_out/scale_r2_n$n.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
d=json.loads(line); print('N=$n', d['value'], 'samples/s', d['ms_per_step'], 'ms/step; e2e', d['e2e']['value'], d['clocks'])"
  fi
done

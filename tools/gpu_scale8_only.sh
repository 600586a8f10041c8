#!/bin/bash
# N = 8 only (after a change that affects the data-parallel run): value, e2e
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/scale_r2b_n8.json 2> gpurun_out/scale_r2b_n8.err
echo "N=8 exit $?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/scale_r2b_n1.json 2> gpurun_out/scale_r2b_n1.err
for n in 8 1; do python -c "
import json
txt=open('gpurun_out/scale_r2b_n$n.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
d=json.loads(line); print('N=$n', d['value'], 'samples/s', d['ms_per_step'], 'ms/step; e2e', d['e2e']['value'], d['clocks'])"; done

#!/bin/bash
# weak-scaling check: bench.py at N = 1, 2, 4, 8 (as many as the box has), one JSON line each
NMAX=$(python -c "import torch; print(torch.cuda.device_count())")
for n in 1 2 4 8; do
  [ $n -gt $NMAX ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 300 --warmup 20 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_$n.json
  fi
  python -c "import json; d=json.load(open('gpurun_out/scale_$n.json')); print('N=$n value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done

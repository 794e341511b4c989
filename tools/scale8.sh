#!/bin/bash
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 300 --warmup 20 2>&1 | grep '^{' | tail -1 > gpurun_out/scale8_diag.json
python -c "import json; d=json.load(open('gpurun_out/scale8_diag.json')); print('N=8 value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['host'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --workload c1 --steps 300 --warmup 20 2>&1 | grep '^{' | tail -1 > gpurun_out/scale8_c1.json
python -c "import json; d=json.load(open('gpurun_out/scale8_c1.json')); print('N=8 c1 value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['host'])"

#!/bin/bash
# weak-scaling check: bench.py at N = 1, 2, 4, 8 (as many as the box has), one JSON line each
O=${1:-gpurun_out/scale}; mkdir -p $O
NMAX=$(python -c "import torch; print(torch.cuda.device_count())")
nproc; nvidia-smi topo -m 2>/dev/null | head -14 > $O/topo.txt
for n in 1 2 4 8; do
  [ $n -gt $NMAX ] && break
  if [ $n -eq 1 ]; then
    timeout -s KILL 400 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu 2>$O/err_$n.txt | tail -1 > $O/scale_$n.json
  else
    timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 300 --warmup 20 2>$O/err_$n.txt | grep '^{' | tail -1 > $O/scale_$n.json
  fi
  python - <<PY
import json
try:
    d=json.load(open('$O/scale_$n.json'))
    print('N=$n value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'host enq', round(d['host']['enqueue_ms_per_step_max_over_ranks'],4), 'e2e', round(d['e2e']['value']), 'h2d GB/s/gpu', round(d['e2e'].get('h2d_gbs_per_gpu',0),1), 'e2e16', round(d['e2e_pcm16']['value']), 'c1 us', round(d['also']['c1']['us_per_launch'],2), 'c4', round(d['also']['c4']['value']), d['also']['c4'].get('allreduce_us'), d['also']['c4'].get('ranks_agree_bitwise'))
except Exception as e:
    print('N=$n FAILED', e); print(open('$O/err_$n.txt').read()[-1500:])
PY
done

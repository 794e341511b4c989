python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 300 --warmup 20 2>&1 | grep '^{' | tail -1 > gpurun_out/final/scale2_c2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --workload c4 --steps 50 --warmup 5 2>&1 | grep '^{' | tail -1 > gpurun_out/final/scale2_c4.json
for f in scale2_c2 scale2_c4; do python -c "import json; d=json.load(open('gpurun_out/final/$f.json')); print('$f N=2 value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['host'])"; done

#!/bin/bash
O=gpurun_out/r2s16; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rs 2>&1 | tail -6 | tee $O/pytest_dist_2gpu.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 2500 $O/bench_n2.json; tail -3 $O/bench_n2.err

#!/bin/bash
O=gpurun_out/r2dist; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rs 2>&1 | tail -6 | tee $O/pytest_dist_2gpu.log

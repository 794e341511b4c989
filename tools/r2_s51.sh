#!/bin/bash
O=gpurun_out/r2s51; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $O/pytest.log
python tools/warpgap.py 2>&1 | tail -10 | tee $O/warp.txt

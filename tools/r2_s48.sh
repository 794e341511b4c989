#!/bin/bash
O=gpurun_out/r2s48; mkdir -p $O
python tools/warpgap.py 2>&1 | tail -6 | tee $O/warp.txt

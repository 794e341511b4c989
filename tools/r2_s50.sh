#!/bin/bash
O=gpurun_out/r2s50; mkdir -p $O
python tools/warpgap.py 2>&1 | tail -5 | tee $O/warp.txt
SFE_LIB_PATH=$PWD/tools/ab/libsfe_base.so python tools/warpgap.py 2>&1 | tail -5 | tee $O/warp_base.txt

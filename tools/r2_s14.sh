#!/bin/bash
O=gpurun_out/r2s14; mkdir -p $O
export SFE_KERNEL=tile
SFE_LIB_PATH=$PWD/tools/ab/libsfe_prof.so timeout -s KILL 120 python tools/prof_tile.py c1 > $O/prof_c1.txt 2>&1
cat $O/prof_c1.txt | cut -c1-200

#!/bin/bash
O=gpurun_out/r2s8; mkdir -p $O
export SFE_KERNEL=tile
SFE_LIB_PATH=$PWD/tools/ab/libsfe_trace.so timeout -s KILL 120 python tools/trace_tile.py c1 > $O/trace_c1.txt 2>&1
SFE_LIB_PATH=$PWD/tools/ab/libsfe_trace_a2.so timeout -s KILL 120 python tools/trace_tile.py c1 > $O/trace_c1_a2.txt 2>&1
head -3 $O/trace_c1.txt; head -3 $O/trace_c1_a2.txt

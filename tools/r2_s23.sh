#!/bin/bash
O=gpurun_out/r2s23; mkdir -p $O
export SFE_LIB_PATH=$PWD/tools/ab/libsfe_trace.so
timeout -s KILL 200 python tools/trace_tile.py c1 ev > $O/trace_c1.txt 2>&1; head -12 $O/trace_c1.txt
timeout -s KILL 200 python tools/trace_tile.py c2 ev > $O/trace_c2.txt 2>&1; head -12 $O/trace_c2.txt

#!/bin/bash
O=gpurun_out/r2s9; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 > $O/pytest.log; cat $O/pytest.log
export SFE_KERNEL=tile
REPS=1 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_nf21.so tools/ab/libsfe_a2.so 2>&1 | tee $O/abn.log
SFE_LIB_PATH=$PWD/tools/ab/libsfe_trace.so timeout -s KILL 120 python tools/trace_tile.py c1 > $O/trace_c1.txt 2>&1
head -2 $O/trace_c1.txt

#!/bin/bash
O=gpurun_out/r2s19; mkdir -p $O
export SFE_KERNEL=tile
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or full_size or coexist or larger_plan" 2>&1 | tail -2
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_fence.so tools/ab/libsfe_nf18.so tools/ab/libsfe_nf22.so 2>&1 | tee $O/abn.log
SFE_LIB_PATH=$PWD/tools/ab/libsfe_prof.so timeout -s KILL 120 python tools/prof_tile.py c1 > $O/prof_c1.txt 2>&1
head -6 $O/prof_c1.txt | cut -c1-170; tail -3 $O/prof_c1.txt | cut -c1-170

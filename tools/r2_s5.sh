#!/bin/bash
O=gpurun_out/r2s5; mkdir -p $O
export SFE_KERNEL=tile
REPS=1 bash tools/abn.sh "c1" speechain_b200/csrc/libsfe.so tools/ab/libsfe_abl1.so tools/ab/libsfe_abl2.so tools/ab/libsfe_abl3.so 2>&1 | tee $O/abn.log

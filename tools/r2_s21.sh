#!/bin/bash
O=gpurun_out/r2s21; mkdir -p $O
export SFE_KERNEL=tile
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_nf13.so tools/ab/libsfe_tw17.so tools/ab/libsfe_tw13.so tools/ab/libsfe_twp13.so 2>&1 | tee $O/abn.log

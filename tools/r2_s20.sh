#!/bin/bash
O=gpurun_out/r2s20; mkdir -p $O
export SFE_KERNEL=tile
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_melsmsp.so tools/ab/libsfe_nf14.so tools/ab/libsfe_nf13r.so tools/ab/libsfe_nf12.so 2>&1 | tee $O/abn.log

#!/bin/bash
O=gpurun_out/r2s39; mkdir -p $O
REPS=2 bash tools/abn.sh "c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_noatom.so 2>&1 | tee $O/abn.log

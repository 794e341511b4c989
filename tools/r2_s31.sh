#!/bin/bash
O=gpurun_out/r2s31; mkdir -p $O
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_npb2nf16.so tools/ab/libsfe_npb2nf18.so tools/ab/libsfe_npb2nf20.so 2>&1 | tee $O/abn.log

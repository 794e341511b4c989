#!/bin/bash
O=gpurun_out/r2s10; mkdir -p $O
export SFE_KERNEL=tile
REPS=1 bash tools/abn.sh "c1" tools/ab/libsfe_a1nf8.so tools/ab/libsfe_a1nf12.so tools/ab/libsfe_a1nf16.so tools/ab/libsfe_a1nf17.so tools/ab/libsfe_a1nf20.so 2>&1 | tee $O/abn.log

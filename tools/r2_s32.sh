#!/bin/bash
O=gpurun_out/r2s32; mkdir -p $O
SFE_LIB_PATH=$PWD/tools/ab/libsfe_npb4nf12.so python bench.py --workload c1 --steps 30 --warmup 5 --no-cpu --no-also 2>&1 | tail -3 | cut -c1-600
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_npb4nf12.so tools/ab/libsfe_npb4nf11.so 2>&1 | grep -v "^  File\|^    \|Traceback\|json" | tee $O/abn.log

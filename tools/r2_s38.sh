#!/bin/bash
O=gpurun_out/r2s38; mkdir -p $O
REPS=2 bash tools/abn.sh "c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_noatom.so 2>&1 | tee $O/abn.log
for lib in speechain_b200/csrc/libsfe.so tools/ab/libsfe_p11nf9.so tools/ab/libsfe_p11nf8.so; do echo "== $lib"; SFE_LIB_PATH=$PWD/$lib python tools/cfg400.py 2>&1 | tail -4; done | tee $O/cfg400.txt
echo "== gather"; SFE_KERNEL=gather python tools/cfg400.py 2>&1 | tail -4 | tee -a $O/cfg400.txt
echo "== gather c2gap"; SFE_KERNEL=gather python tools/c2gap.py 2>&1 | grep ragged | tee $O/c2gap_gather.txt

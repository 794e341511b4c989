#!/bin/bash
O=gpurun_out/r2s52; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
for lib in tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so; do echo "== $lib"; SFE_LIB_PATH=$PWD/$lib python tools/fancygap.py 2>&1 | tail -14 | head -4; SFE_LIB_PATH=$PWD/$lib python tools/c2gap.py 2>&1 | grep ragged; done 2>&1 | tee $O/gap.txt
REPS=2 bash tools/abn.sh "c1 c2" tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so 2>&1 | tee $O/abn.log

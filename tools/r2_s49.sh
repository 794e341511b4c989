#!/bin/bash
O=gpurun_out/r2s49; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
python tools/warpgap.py 2>&1 | tail -6 | tee $O/warp.txt
REPS=1 bash tools/abn.sh "c1 c2" tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so 2>&1 | tee $O/abn.log

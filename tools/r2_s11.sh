#!/bin/bash
O=gpurun_out/r2s11; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
export SFE_KERNEL=tile
REPS=1 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_a1.so tools/ab/libsfe_a2.so 2>&1 | tee $O/abn.log

#!/bin/bash
O=gpurun_out/r2s33; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
REPS=3 bash tools/abn.sh "c1 c2" tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so 2>&1 | tee $O/abn.log

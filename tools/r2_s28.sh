#!/bin/bash
O=gpurun_out/r2s28; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_ms2nf14.so tools/ab/libsfe_nf14.so tools/ab/libsfe_nf15.so tools/ab/libsfe_nf16.so 2>&1 | tee $O/abn.log

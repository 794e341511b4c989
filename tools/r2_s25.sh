#!/bin/bash
O=gpurun_out/r2s25; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
REPS=2 bash tools/abenv.sh "c1 c2" SFE_HOST_SPLIT=0 SFE_HOST_SPLIT=1 2>&1 | tee $O/ab.log

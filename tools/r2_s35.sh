#!/bin/bash
O=gpurun_out/r2s35; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log

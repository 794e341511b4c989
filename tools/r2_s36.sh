#!/bin/bash
O=gpurun_out/r2s36; mkdir -p $O
python tools/c2gap.py 2>&1 | tee $O/c2gap.txt

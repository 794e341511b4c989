#!/bin/bash
O=gpurun_out/r2s24; mkdir -p $O
timeout -s KILL 300 python tools/fixedcost.py c1 2>&1 | tee $O/fixed_c1.txt
timeout -s KILL 300 python tools/fixedcost.py c2 2>&1 | tee $O/fixed_c2.txt

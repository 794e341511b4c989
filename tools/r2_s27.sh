#!/bin/bash
O=gpurun_out/r2s27; mkdir -p $O
timeout -s KILL 300 python tools/fixedcost.py c1 2>&1 | tee $O/fixed_c1.txt
timeout -s KILL 300 python tools/hostbreak.py 2>&1 | tee $O/hostbreak.txt
nproc; lscpu | grep -i "model name\|^CPU(s)\|MHz" | head -5

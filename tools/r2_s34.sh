#!/bin/bash
O=gpurun_out/r2s34; mkdir -p $O
python tools/h2d_probe2.py 2>&1 | tee $O/h2d.txt
for wl in c1; do python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl e2e', d['e2e'])"; done

#!/bin/bash
# A/B: same box, alternate base/new builds; prints kernel us and ms/step per workload
WLS=${1:-"c2 c1"}
for rep in 1 2; do
for wl in $WLS; do
for lib in tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so; do
SFE_LIB_PATH=$PWD/$lib python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', '$lib'.split('/')[-1], 'ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2))"
done; done; done

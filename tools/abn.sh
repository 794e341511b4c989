#!/bin/bash
# A/B/n on one box: [REPS=2] tools/abn.sh "c2 c1" libA.so libB.so ...  (alternating)
WLS=$1; shift
for rep in $(seq 1 ${REPS:-2}); do
for wl in $WLS; do
for lib in "$@"; do
SFE_LIB_PATH=$PWD/$lib python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', '$lib'.split('/')[-1], 'ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2))"
done; done; done

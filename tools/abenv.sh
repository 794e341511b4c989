#!/bin/bash
# A/B of environment switches on one box: [REPS=2] tools/abenv.sh "c1 c2" "VAR=a" "VAR=b" ...   (alternating)
WLS=$1; shift
for rep in $(seq 1 ${REPS:-2}); do
for wl in $WLS; do
for ev in "$@"; do
env $ev python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu --no-also 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', '$ev', 'ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2), 'host', round(d['host']['enqueue_ms_per_step_max_over_ranks'],4))"
done; done; done

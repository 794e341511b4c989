#!/bin/bash
O=gpurun_out/r2s15; mkdir -p $O
for wl in c3 c4 c5; do for k in gather tile; do
SFE_KERNEL=$k timeout -s KILL 200 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu 2>$O/err_${wl}_$k.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl $k', 'ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],3))" || tail -3 $O/err_${wl}_$k.txt
done; done

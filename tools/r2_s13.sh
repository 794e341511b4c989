#!/bin/bash
O=gpurun_out/r2s13; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for wl in c1 c2; do for k in gather tile; do
SFE_KERNEL=$k python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl $k', 'ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2))"
done; done
export SFE_KERNEL=tile
SFE_LIB_PATH=$PWD/tools/ab/libsfe_prof.so timeout -s KILL 120 python tools/prof_tile.py c1 > $O/prof_c1.txt 2>&1
head -8 $O/prof_c1.txt; tail -3 $O/prof_c1.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_c1_tile python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu > $O/ncu_c1.log 2>&1

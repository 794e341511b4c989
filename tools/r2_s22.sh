#!/bin/bash
# ncu evidence for the tile kernel: launch list (C2 step) + full captures of the tile kernel on C1 and C2 and of the normalise pass
O=gpurun_out/r2s22; mkdir -p $O
timeout -s KILL 300 python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.json; echo
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-also > $O/ncu_launch.log 2>&1
for wl in c1 c2; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_${wl}_tile python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-also > $O/ncu_$wl.log 2>&1
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:normalize_kernel --launch-skip 6 --launch-count 1 -f -o $O/full_c2_normalize python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-also > $O/ncu_c2n.log 2>&1
ls -la $O

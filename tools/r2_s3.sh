#!/bin/bash
O=gpurun_out/r2s3; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log; cat $O/pytest.log
for wl in c1 c2; do
  for k in gather tile; do
    SFE_KERNEL=$k timeout -s KILL 200 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu > $O/bench_${wl}_$k.json 2> $O/bench_${wl}_$k.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${wl}_$k.json").read().strip().splitlines()[-1])
    print("$wl $k ms/step", round(d["ms_per_step"],4), "kernel us", round(d["roofline"]["us_per_launch"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$wl $k FAILED", e); print(open("$O/bench_${wl}_$k.err").read()[-1500:])
PY
  done
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_c1_tile python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu > $O/ncu_c1.log 2>&1

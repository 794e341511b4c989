#!/bin/bash
# Round-end measurement on one GPU: tests, bench lines for every workload, reference arm, launch list and full ncu
# captures (every ncu command only after the same command has exited 0 without ncu).
O=gpurun_out/r2final; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest.log; cat $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
for wl in c2 c1 c3 c4 c5; do
  timeout -s KILL 400 python bench.py --workload $wl > $O/bench_$wl.json 2> $O/bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "kernel us", round(d["roofline"]["us_per_launch"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), "cpu", round(d.get("cpu_baseline",{}).get("value",0)), d.get("cpu_baseline",{}).get("kind"))
except Exception as e:
    print("$wl FAILED", e); print(open("$O/bench_$wl.err").read()[-800:])
PY
done
timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 700 $O/bench_ref.json
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-also > /dev/null 2>&1 && \
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-also > $O/ncu_launch.log 2>&1
for wl in c1 c2; do
  timeout -s KILL 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-also > /dev/null 2>&1 && \
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_${wl}_tile python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-also > $O/ncu_$wl.log 2>&1
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:normalize_kernel --launch-skip 6 --launch-count 1 -f -o $O/full_c2_normalize python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu --no-also > $O/ncu_c2n.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:fused_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_c3_fused python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-also > $O/ncu_c3.log 2>&1
ls -la $O

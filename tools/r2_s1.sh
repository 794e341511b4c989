#!/bin/bash
# round 2, first run of the tile kernel: targeted parity, full suite, A/B against the gather kernel
O=gpurun_out/r2s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "speech2mel_golden" 2>&1 | tail -15 > $O/pytest_quick.log; cat $O/pytest_quick.log
if grep -q "passed" $O/pytest_quick.log && ! grep -q "failed" $O/pytest_quick.log; then
  timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.log; cat $O/pytest.log
fi
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

#!/bin/bash
O=gpurun_out/r2s17; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for pdl in 0 1; do
SFE_PDL=$pdl timeout -s KILL 300 python bench.py --steps 300 --warmup 20 --no-cpu > $O/bench_c2_pdl$pdl.json 2> $O/err_pdl$pdl.txt
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_c2_pdl$pdl.json").read().strip().splitlines()[-1])
    print("pdl=$pdl c2 ms/step", round(d["ms_per_step"],4), "kernel us", round(d["roofline"]["us_per_launch"],2), "host enq", round(d["host"]["enqueue_ms_per_step_max_over_ranks"],4), "e2e", round(d["e2e"]["value"]), "e2e16", round(d["e2e_pcm16"]["value"]), "c1 us", round(d["also"]["c1"]["us_per_launch"],2), "c4 ms", round(d["also"]["c4"]["ms_per_step"],4))
except Exception as e:
    print("FAILED", e); print(open("$O/err_pdl$pdl.txt").read()[-1500:])
PY
done

#!/bin/bash
O=gpurun_out/r2s29; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_nf17.so 2>&1 | tee $O/abn.log
timeout -s KILL 300 python tools/hostbreak.py 2>&1 | tee $O/hostbreak.txt
timeout -s KILL 300 python bench.py --steps 300 --warmup 20 --no-cpu > $O/bench_c2.json 2> $O/bench_c2.err; python - <<PY
import json
d=json.loads(open("$O/bench_c2.json").read().strip().splitlines()[-1])
print("c2 ms/step", round(d["ms_per_step"],4), "kernel us", round(d["roofline"]["us_per_launch"],2), "frac", round(d["roofline"]["frac"],3), "host enq", round(d["host"]["enqueue_ms_per_step_max_over_ranks"],4), "e2e", round(d["e2e"]["value"]), "e2e16", round(d["e2e_pcm16"]["value"]), "c1 us", round(d["also"]["c1"]["us_per_launch"],2), "c4 ms", round(d["also"]["c4"]["ms_per_step"],4))
PY

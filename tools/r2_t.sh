#!/bin/bash
ls -la speechain_b200/csrc/drawext/
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout -s KILL 300 python tools/hostbreak.py 2>&1 | tail -12
python tools/warpgap.py 2>&1 | sed -n 1,5p
for v in 1 0; do SFE_DRAW_HELPER=$v python bench.py --steps 300 --warmup 20 --no-cpu --no-also 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('helper=$v c2 ms/step', round(d['ms_per_step'],4), 'host', round(d['host']['enqueue_ms_per_step_max_over_ranks'],4))"; done

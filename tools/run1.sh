timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in c4 c5 c2; do python bench.py --workload $wl --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2), 'value', round(d['value']))"; done

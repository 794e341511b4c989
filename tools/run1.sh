timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for k in 2 4 8; do echo "norm ctas/sm $k"; SFE_NORM_CTAS_PER_SM=$k python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'kernel us', round(d['roofline']['us_per_launch'],2))"; done

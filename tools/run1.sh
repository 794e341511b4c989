timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ragged_upload or native" 2>&1 | tail -2
for g in 0 1 0 1; do SFE_UPLOAD_GROUP=$g python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('group $g e2e', round(d['e2e']['value']), 'h2d', d['e2e']['h2d_bytes_per_step'])"; done

#!/bin/bash
O=gpurun_out/r2s30; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/pytest.log
REPS=2 bash tools/abn.sh "c2" tools/ab/libsfe_base.so speechain_b200/csrc/libsfe.so 2>&1 | tee $O/abn.log
for v in 3 5 6; do
SFE_NORM_CTAS_PER_SM=$v python bench.py --workload c2 --steps 300 --warmup 20 --no-cpu --no-also 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('per_sm $v ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'plain h2d', round(d['e2e']['h2d_plain_copy_gbs_per_gpu'],1), 'e2e16', round(d['e2e_pcm16']['value']))"
done 2>&1 | tee -a $O/abn.log

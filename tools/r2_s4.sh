#!/bin/bash
O=gpurun_out/r2s4; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/pytest.log; cat $O/pytest.log
export SFE_KERNEL=tile
REPS=2 bash tools/abn.sh "c1 c2" speechain_b200/csrc/libsfe.so tools/ab/libsfe_nf21.so tools/ab/libsfe_nf17.so 2>&1 | tee $O/abn.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_c1_tile python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu > $O/ncu_c1.log 2>&1

#!/bin/bash
# Round-end measurement on one GPU: tests, bench lines for every workload, launch list and full ncu captures.
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/final/pytest.log; cat gpurun_out/final/pytest.log
for wl in c2 c1 c3 c4 c5; do
  timeout 300 python bench.py --workload $wl > gpurun_out/final/bench_$wl.json 2> gpurun_out/final/bench_$wl.err
  tail -c 600 gpurun_out/final/bench_$wl.json | head -c 300; echo
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/final/launches_c2.csv python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/final/ncu_launch.log 2>&1
for wl in c1 c2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_frontend --launch-skip 6 --launch-count 1 -f -o gpurun_out/final/full_${wl}_fused python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/final/ncu_$wl.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:normalize_kernel --launch-skip 6 --launch-count 1 -f -o gpurun_out/final/full_c2_normalize python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/final/ncu_c2n.log 2>&1
ls -la gpurun_out/final

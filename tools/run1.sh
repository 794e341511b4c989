mkdir -p gpurun_out/final
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2
for wl in c2 c1 c3 c4 c5; do
  timeout 200 python bench.py --workload $wl > gpurun_out/final/bench_$wl.json 2> gpurun_out/final/bench_$wl.err
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

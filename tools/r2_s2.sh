#!/bin/bash
O=gpurun_out/r2s2; mkdir -p $O
for t in test_c3_full_size_vs_oracle test_repeated_calls_are_deterministic_and_scheduler_rearms "test_work_distribution_edge_batches" test_memory_beyond_each_length_is_never_read test_ragged_upload_equals_padded_copy test_fused_chain_with_time_warp_vs_oracle test_c5_style_linear_and_mel_from_one_launch_with_running_global_norm test_checkpoint_load_rebuilds_the_plan; do
  timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$t" 2>&1 | tail -25 > $O/t_$t.log
  echo "== $t: $(tail -1 $O/t_$t.log)"
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tile_frontend --launch-skip 6 --launch-count 1 -f -o $O/full_c1_tile python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu > $O/ncu_c1.log 2>&1
tail -3 $O/ncu_c1.log
ls -la $O

#!/bin/bash
# host-side cost of one chain step with N concurrent processes (CPU contention check)
nproc; python -c "import os; print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))"
for n in 1 4 8; do
  echo "--- $n concurrent processes"
  for i in $(seq 1 $n); do (OMP_NUM_THREADS=1 python tools/hostprof.py 2>&1 | grep "host per step" &) ; done
  sleep 25
done

#!/bin/bash
# gpurun_out/r2final (written by tools/final_run.sh on the GPU box) -> profiles/r2_* (run here, where ncu reads the reports)
set -e
cd "$(dirname "$0")/.."
R=gpurun_out/r2final; B=speechain_b200/csrc/build
K1='tile_frontend_kernelINS_7FftPlanILi256ELi16ELi16ELi16ELi1ELi16ELi400EEELb0ELb0'
K2='tile_frontend_kernelINS_7FftPlanILi256ELi16ELi16ELi16ELi1ELi16ELi400EEELb1ELb0'
K3='fused_frontend_kernelINS_7FftPlanILi512ELi32ELi4ELi8ELi16ELi16ELi1024EEELb0EEE'
(python profiles/ncu_summary.py $R/full_c1_tile.ncu-rep; python profiles/ncu_by_line.py $R/full_c1_tile.ncu-rep $B/sfe_plan_12.o $K1 40) > profiles/r2_ncu_full_c1_tile.txt 2>&1
(python profiles/ncu_summary.py $R/full_c2_tile.ncu-rep; python profiles/ncu_by_line.py $R/full_c2_tile.ncu-rep $B/sfe_plan_12.o $K2 40) > profiles/r2_ncu_full_c2_tile.txt 2>&1
(python profiles/ncu_summary.py $R/full_c2_normalize.ncu-rep; python profiles/ncu_by_line.py $R/full_c2_normalize.ncu-rep $B/sfe_lib.o normalize_kernelILb0 40) > profiles/r2_ncu_full_c2_normalize.txt 2>&1
(python profiles/ncu_summary.py $R/full_c3_fused.ncu-rep; python profiles/ncu_by_line.py $R/full_c3_fused.ncu-rep $B/sfe_plan_13.o $K3 40) > profiles/r2_ncu_full_c3_fused.txt 2>&1
python profiles/ncu_hot_sass.py $R/full_c1_tile.ncu-rep $B/sfe_plan_12.o $K1 24032 0.5 > profiles/r2_sass_hot_c1_tile.txt 2>&1
for wl in c1 c2 c3 c4 c5; do cp $R/bench_$wl.json profiles/r2_bench_$wl.json; done
cp $R/bench_ref.json profiles/r2_bench_ref_c2.json; cp $R/launches_c2.csv profiles/r2_launches_c2.csv
cat $R/pytest.log $R/smoke.log > profiles/r2_pytest_gpu.log
python - <<'PY'
import json, re
m = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}
for wl, f in (('c1', 'profiles/r2_ncu_full_c1_tile.txt'), ('c2', 'profiles/r2_ncu_full_c2_tile.txt'), ('c3', 'profiles/r2_ncu_full_c3_fused.txt')):
    t = open(f).read()
    r = re.search(r'dram__bytes_read.sum\s+([\d.]+)\s+(\w+)', t); w = re.search(r'dram__bytes_write.sum\s+([\d.]+)\s+(\w+)', t)
    p = f'profiles/traffic_{wl}.json'; d = json.load(open(p))
    d['dram_bytes_read'] = int(round(float(r.group(1)) * m[r.group(2)])); d['dram_bytes_write'] = int(round(float(w.group(1)) * m[w.group(2)]))
    d['dram_bytes_per_launch'] = d['dram_bytes_read'] + d['dram_bytes_write']
    json.dump(d, open(p, 'w'), indent=1)
    print(wl, d['dram_bytes_read'], d['dram_bytes_write'])
PY
grep -n "sass instrs" profiles/r2_ncu_full_*.txt

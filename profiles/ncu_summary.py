#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics, stall reasons, opcode mix (run where ncu is installed).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + which]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
print("kernel:", data[hdr.index("Kernel Name")][:100])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"  {w:72s} {data[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur, h = [], [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        if cur:
            blocks.append(cur)
        cur = []
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if r:
        cur.append(r)
if cur:
    blocks.append(cur)
b = blocks[which]
idx = {k: i for i, k in enumerate(h)}
stall = collections.Counter()
for r in b:
    for k in h:
        if k.startswith("stall_") and "Not Issued" not in k:
            stall[k] += int(r[idx[k]] or 0)
S = sum(stall.values())
print("stall reasons (all samples):")
for k, v in stall.most_common(10):
    print(f"  {k:28s} {100 * v / S:5.1f}%")
ops, samp = collections.Counter(), collections.Counter()
for r in b:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[idx["Instructions Executed"]] or 0)
    samp[op] += int(r[idx["# Samples"]] or 0)
T = sum(ops.values())
print(f"opcode mix ({T} warp instructions, {len(b)} SASS lines):")
for op, v in ops.most_common(22):
    print(f"  {op:10s} {100 * v / T:5.1f}%   stall samples {100 * samp[op] / max(1, sum(samp.values())):5.1f}%")

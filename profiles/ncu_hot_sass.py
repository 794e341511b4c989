#!/usr/bin/env python
"""Annotated SASS of a kernel's hot path from an ncu --set full --import-source on capture: every instruction that
was executed at least `frac` x (the executions of the median hot instruction), in address order, with its source
line (nvdisasm --print-line-info of the object file), executions per warp-level unit and stall samples.
usage: python profiles/ncu_hot_sass.py <rep> <plan object .o> <kernel substring> [units (e.g. pairs per launch)] [frac]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    if ksub in out:
        dis = out.splitlines()
        break
lines, cur_line, infn = [], None, False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur_line)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, b = None, []
for r in rows:
    if r and r[0] == "Address":
        h = r
        continue
    if r and r[0] == "Kernel Name":
        continue
    if r and h:
        b.append(r)
idx = {k: i for i, k in enumerate(h)}
ex = [int(r[idx["Instructions Executed"]] or 0) for r in b]
sm = [int(r[idx["# Samples"]] or 0) for r in b]
# the hot level: the execution count shared by the largest number of instructions
lvl = collections.Counter(e for e in ex if e > 0)
hot = max(lvl.items(), key=lambda kv: kv[0] * kv[1])[0]
print(f"# {ksub}: {len(b)} SASS instructions, {sum(ex)} executed (warp level), {sum(sm)} stall samples")
print(f"# hot level: {hot} executions ({lvl[hot]} instructions at exactly that count); listed: executions >= {frac} x hot")
if units:
    print(f"# units per launch {units:.0f}: hot level = {hot / units:.2f} executions per unit")
base = int(b[0][idx["Address"]], 16)
ops = collections.Counter()
tot = 0
print(f"# {'offset':>7s} {'exec':>9s} {'samples':>7s}  {'source line':24s} instruction")
for i, r in enumerate(b):
    if ex[i] < frac * hot:
        continue
    a = int(r[idx["Address"]], 16) - base
    t = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t)
    ops[m.group(2).split(".")[0] if m else "?"] += ex[i]
    tot += ex[i]
    print(f"  {a:7x} {ex[i]:9d} {sm[i]:7d}  {str(lines[i] if i < len(lines) else '?'):24s} {t}")
print("# opcode mix of the listed instructions (share of their executions):")
for k, v in ops.most_common(30):
    print(f"#   {k:10s} {100 * v / tot:5.1f}%" + (f"  {v / units:7.1f} per unit" if units else ""))

#!/usr/bin/env python
"""Join ncu's per-SASS-instruction counters with nvdisasm line info -> per source line totals.
usage: python profiles/ncu_by_line.py <rep> <libsfe.so> <kernel substring> [top N]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):  # one cubin per translation unit
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    if ksub in out:
        dis = out.splitlines()
        break
# collect (instruction order -> source line) for the wanted function
lines, cur_line, infn = [], None, False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur_line)
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
b = blocks[0]
idx = {k: i for i, k in enumerate(h)}
print(f"sass instrs: ncu {len(b)}  nvdisasm {len(lines)}")
n = min(len(b), len(lines))
ex, st = collections.Counter(), collections.Counter()
for i in range(n):
    ex[lines[i]] += int(b[i][idx["Instructions Executed"]] or 0)
    st[lines[i]] += int(b[i][idx["# Samples"]] or 0)
T, S = sum(ex.values()), sum(st.values())
print(f"{'file:line':28s} {'inst%':>6s} {'stall%':>6s}")
for k, v in ex.most_common(top):
    print(f"{str(k[0]) + ':' + str(k[1]) if k else '?':28s} {100 * v / T:6.2f} {100 * st[k] / max(S, 1):6.2f}")

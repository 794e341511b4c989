"""Annotated hot SASS: ncu per-instruction counters joined with nvdisasm line info.
usage: python tools/sass_annot.py <rep> <object or .so with ONE cubin holding the kernel> <kernel substring> > out.txt"""
import csv, glob, io, os, re, subprocess, sys, tempfile
rep, so, ksub = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    if ksub in out:
        dis = out.splitlines(); break
lines, cur, infn = [], None, False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infn = ksub in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = f"{os.path.basename(m.group(1)).replace('sfe_','')}:{m.group(2)}"
        if "inlined at" in ln:
            pass
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; body = []
for r in rows:
    if r and r[0] == "Address": h = r; continue
    if r and r[0] == "Kernel Name": continue
    if r and h: body.append(r)
idx = {k: i for i, k in enumerate(h)}
tot = sum(int(r[idx["Instructions Executed"]] or 0) for r in body)
ts = sum(int(r[idx["# Samples"]] or 0) for r in body)
print(f"# total warp instrs {tot}  samples {ts}")
for i, r in enumerate(body):
    ex = int(r[idx["Instructions Executed"]] or 0); sm = int(r[idx["# Samples"]] or 0)
    print(f"{i:6d} {ex:9d} {100*ex/tot:5.2f}% s{100*sm/ts:05.2f}% {lines[i] if i < len(lines) else '?':18s} {r[idx['Source']]}")

"""Executed warp-instructions per REGION of the fused kernel: ncu per-SASS counters joined with the OUTERMOST
source line (nvdisasm -gi inline chains), bucketed by line ranges of the kernel body.
usage: python tools/region_profile.py <rep> <plan .o/.so> <kernel substring> <units> [regions.txt]
regions file: lines "<first line> <name>" (sorted); default = every 1 line."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, so, ksub, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
regions = []
kfile = os.environ.get("KFILE", "sfe_fused.cuh")
if len(sys.argv) > 5:
    for ln in open(sys.argv[5]):
        a = ln.split(None, 1)
        if a: regions.append((int(a[0]), a[1].strip()))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    if ksub in out:
        dis = out.splitlines(); break
outer, ops, infn, chain = [], [], False, []
fresh = False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        infn = ksub in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if not fresh: chain = []; fresh = True
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3): chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    mi = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if mi:
        fresh = False
        outer.append(chain[-1] if chain else None)
        ops.append(mi.group(1))
if rep == "-":  # static mode: every SASS instruction counts once (straight-line regions: static == executed per unit)
    h = ["Instructions Executed", "# Samples"]
    body = [["1", "0"] for _ in outer]
    units = 1.0
else:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, body = None, []
    for r in rows:
        if r and r[0] == "Address": h = r; continue
        if r and r[0] == "Kernel Name":
            if body: break
            continue
        if r and h: body.append(r)
idx = {k: i for i, k in enumerate(h)}
n = min(len(body), len(outer))
print(f"# sass: ncu {len(body)} nvdisasm {len(outer)}")
def reg(loc):
    if loc is None: return "?"
    f, l = loc
    if not regions: return f"{f}:{l}"
    if f != kfile: return f
    name = "pre"
    for a, nm in regions:
        if l >= a: name = nm
    return name
ex, st, opc = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for i in range(n):
    e = int(body[i][idx["Instructions Executed"]] or 0); s = int(body[i][idx["# Samples"]] or 0)
    r = reg(outer[i]); ex[r] += e; st[r] += s
    o = ops[i].split()
    o = o[1] if o[0].startswith("@") else o[0]
    opc[r][o.split(".")[0]] += e
T, S = sum(ex.values()), sum(st.values())
print(f"# total {T} warp-instr = {T/units:.1f} per unit")
for k, v in sorted(ex.items(), key=lambda kv: -kv[1]):
    top = ", ".join(f"{o} {c/units:.0f}" for o, c in opc[k].most_common(8))
    print(f"{k:28s} {v/units:8.1f}/unit {100*v/T:5.1f}%  stall {100*st[k]/max(S,1):5.1f}%   {top}")

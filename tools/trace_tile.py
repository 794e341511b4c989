"""Development: timeline of the tile kernel's CTA 0 and the entry/exit times of every warp of every CTA, from a
-DSFE_T2_TRACE build.
    python tools/variant.py trace "-DSFE_T2_TRACE" 12 ; SFE_LIB_PATH=tools/ab/libsfe_trace.so python tools/trace_tile.py [c1|c2] [events]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.module.acoustic_frontend import AcousticFrontend
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
dev = torch.device("cuda", 0)
lens = bench.workload_lengths(w)
chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=w["normalize"], specaug=w["specaug"]).to(dev)
chain.train()
wavs = [bench.make_batch(w, lens, i, dev) for i in range(4)]
wl = torch.tensor(lens, dtype=torch.int64, device=dev)
lib = C.CDLL(os.environ["SFE_LIB_PATH"])
cap = 1 << 15
t = np.zeros(cap, dtype=np.uint64); tag = np.zeros(cap, dtype=np.uint32)
for it in range(3):  # back-to-back launches like the bench loop; the LAST launch is the one read back
    for k in range(8):
        chain(wavs[k % 4], wl)
    torch.cuda.synchronize()
    n = lib.sfe_debug_trace(C.c_void_p(t.ctypes.data), C.c_void_p(tag.ctypes.data), cap)
span = np.zeros((256, 32, 2), dtype=np.uint64)
lib.sfe_debug_span(C.c_void_p(span.ctypes.data))
ev = sorted((int(t[i]), int(tag[i]) >> 24, (int(tag[i]) >> 16) & 0xff, int(tag[i]) & 0xffff) for i in range(n))
# the trace buffer holds the 8 launches of the last round: keep the last launch (events after the last tag-1 burst)
starts = [e[0] for e in ev if e[1] == 1 and e[2] == 0]
last0 = starts[-1] if starts else ev[0][0]
ev = [e for e in ev if e[0] >= last0 - 3000]
used = span[:, :, 0] > 0
ent = span[:, :, 0].astype(np.int64); ext = span[:, :, 1].astype(np.int64)
g0 = ent[used].min()
print("grid: CTAs", int(used.any(1).sum()), "warps/CTA", int(used[0].sum()))
print("  kernel span (first entry -> last exit) us", (ext[used].max() - g0) / 1e3)
ce = np.where(used, ent, 1 << 62).min(1)[used.any(1)]; cx = np.where(used, ext, 0).max(1)[used.any(1)]
print("  CTA first-warp entry  us: min %.2f median %.2f max %.2f" % tuple((np.percentile(ce, q) - g0) / 1e3 for q in (0, 50, 100)))
le = np.where(used, ent, 0).max(1)[used.any(1)]
print("  CTA last-warp entry   us: min %.2f median %.2f max %.2f" % tuple((np.percentile(le, q) - g0) / 1e3 for q in (0, 50, 100)))
print("  CTA last-warp exit    us: min %.2f median %.2f max %.2f" % tuple((np.percentile(cx, q) - g0) / 1e3 for q in (0, 50, 100)))
nf = int(used[0].sum()) - 3
fx = np.where(used, ext, 0)[:, :nf].max(1)[used.any(1)]
print("  CTA last FFT-warp exit us: min %.2f median %.2f max %.2f" % tuple((np.percentile(fx, q) - g0) / 1e3 for q in (0, 50, 100)))
print("  CTA 0: warp entry offsets ns", [int(x - ent[0][used[0]].min()) for x in ent[0][used[0]]])
print("  CTA 0: warp exit offsets (before the CTA's end) ns", [int(ext[0][used[0]].max() - x) for x in ext[0][used[0]]])
names = {1: "entry", 2: "setup_done", 3: "tables_ok", 10: "M.pfull_ok", 11: "M.prog_done", 12: "M.readout_done",
         20: "F.claim", 21: "F.full_ok", 22: "F.pass1_done", 23: "F.unpack_done", 24: "F.p2_ok", 25: "F.published"}
if len(sys.argv) > 2:
    t0 = ev[0][0]
    print("events", len(ev), "span_us", (ev[-1][0] - t0) / 1e3)
    for tm, tg, wp, idx in ev:
        print(f"{(tm - t0) / 1e3:9.3f} us  w{wp:02d} {names.get(tg, tg):16s} {idx}")

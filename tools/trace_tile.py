"""Development: timeline of the tile kernel's CTA 0 from a -DSFE_T2_TRACE build.
    python tools/variant.py trace "-DSFE_T2_TRACE" 12 ; SFE_LIB_PATH=tools/ab/libsfe_trace.so python tools/trace_tile.py [c1|c2]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200 import _lib
from speechain_b200.module.acoustic_frontend import AcousticFrontend
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
dev = torch.device("cuda", 0)
lens = bench.workload_lengths(w)
chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=w["normalize"], specaug=w["specaug"]).to(dev)
chain.train()
wav = bench.make_batch(w, lens, 1, dev)
wl = torch.tensor(lens, dtype=torch.int64, device=dev)
lib = C.CDLL(os.environ["SFE_LIB_PATH"])
cap = 1 << 15
t = np.zeros(cap, dtype=np.uint64); tag = np.zeros(cap, dtype=np.uint32)
for it in range(3):
    chain(wav, wl); torch.cuda.synchronize()
    n = lib.sfe_debug_trace(C.c_void_p(t.ctypes.data), C.c_void_p(tag.ctypes.data), cap)
ev = sorted((int(t[i]), int(tag[i]) >> 24, (int(tag[i]) >> 16) & 0xff, int(tag[i]) & 0xffff) for i in range(n))
t0 = ev[0][0]
names = {1: "P.sfree_ok", 2: "P.issued", 3: "P.described", 10: "M.pfull_ok", 11: "M.prog_done", 12: "M.readout_done",
         20: "F.claim", 21: "F.full_ok", 22: "F.pass1_done", 23: "F.unpack_done", 24: "F.p2_ok", 25: "F.published"}
print("events", n, "span_us", (ev[-1][0] - t0) / 1e3)
for tm, tg, wp, idx in ev:
    print(f"{(tm - t0) / 1e3:9.3f} us  w{wp:02d} {names.get(tg, tg):16s} {idx}")

"""Development: per-warp phase cycle totals of the tile kernel's CTA 0 from a -DSFE_T2_PROF build.
    python tools/variant.py prof "-DSFE_T2_PROF" 12 ; SFE_LIB_PATH=tools/ab/libsfe_prof.so python tools/prof_tile.py [c1|c2]"""
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
wav = bench.make_batch(w, lens, 1, dev)
wl = torch.tensor(lens, dtype=torch.int64, device=dev)
lib = C.CDLL(os.environ["SFE_LIB_PATH"])
out = np.zeros(32 * 16, dtype=np.int64)
for it in range(5):
    chain(wav, wl); torch.cuda.synchronize()
lib.sfe_debug_prof(C.c_void_p(out.ctypes.data))
out = out.reshape(32, 16)
print("FFT warps: cycles in [0 data wait, 1 pass1, 2 claim, 3 lastpass+unpack, 4 p2 wait, 5 store, 6 publish+loop, 7 skipped pairs]")
print("Mel warps: cycles in [0 tile wait, 1 flags, 2 program, 3 read-out, 4 hand-back]")
for wi in range(32):
    if out[wi].any():
        g = out[wi][8:12]
        print(f"w{wi:02d} total {out[wi][:8].sum():8d} ", " ".join(f"{x:7d}" for x in out[wi][:8]),
              f" | ns: setup {g[1] - g[0]:6d} tables {g[2] - g[1]:6d} loop {g[3] - g[2]:6d} all {g[3] - g[0]:6d}")

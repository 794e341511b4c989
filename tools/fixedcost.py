"""Development: fused kernel time against batch size (fixed cost per launch vs steady-state cost per pair).
    [SFE_LIB_PATH=...] python tools/fixedcost.py [c1|c2]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.host import ops
from speechain_b200.module.acoustic_frontend import AcousticFrontend
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
dev = torch.device("cuda", 0)
pts = []
for B in (16, 32, 48, 64, 128, 256):
    w = dict(bench.WORKLOADS[name]); w["B"] = B
    lens = bench.workload_lengths(w)
    chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=None, specaug=None).to(dev)
    fe = chain.frontend
    T = max(lens)
    ring = max(2, int(np.ceil(3 * bench.L2_BYTES / (4 * B * T * 1.33))) + 1)
    batches = [bench.make_batch(w, lens, 7 + i, dev) for i in range(ring)]
    wl = torch.tensor(lens, dtype=torch.int64, device=dev)
    ops.HostLengths.remember(wl, np.asarray(lens, dtype=np.int64))
    for i in range(5):
        fe.run(batches[i % ring], wl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for i in range(n):
        fe.run(batches[i % ring], wl)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    pairs = sum((1 + l // 160 + 1) // 2 for l in lens)
    pts.append((pairs, us))
    print(f"{name} B={B:4d} ring {ring:3d} pairs {pairs:7d}  {us:8.2f} us/launch  {us * 1e3 / pairs * 148:7.1f} ns per pair per SM")
    del batches
x = np.array([p[0] for p in pts], float); y = np.array([p[1] for p in pts], float)
b, a = np.polyfit(x, y, 1)
print(f"fit: {a:.2f} us fixed + {b * 1e3 * 148:.1f} ns per pair per SM")

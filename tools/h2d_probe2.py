"""Probe: the library's upload against plain copies for the C1 (32 x 15 s, equal lengths) and C2 shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speechain_b200.host import ops
def run(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n
for name in ("c1", "c2"):
    w = bench.WORKLOADS[name]
    lens = bench.workload_lengths(w)
    B, T = len(lens), max(lens)
    host = torch.randn(B, T, 1).pin_memory()
    dev = torch.empty(B, T, 1, device="cuda")
    wl = torch.tensor(lens, dtype=torch.int64)
    tot = ops.upload_bytes(wl)
    t = run(lambda: ops.upload_waveforms(host, wl, out=dev)); print(f"{name} upload_waveforms            {t*1e3:7.3f} ms {tot/t/1e9:6.1f} GB/s ({tot/1e6:.1f} MB)")
    def rows():
        for b, n in enumerate(lens): dev[b, :n].copy_(host[b, :n], non_blocking=True)
    t = run(rows); print(f"{name} one copy per row            {t*1e3:7.3f} ms {sum(lens)*4/t/1e9:6.1f} GB/s")
    t = run(lambda: dev.copy_(host, non_blocking=True)); print(f"{name} padded (B,T) copy           {t*1e3:7.3f} ms {B*T*4/t/1e9:6.1f} GB/s (all bytes)")
    for k in (1, 2, 4, 8):
        os.environ["X"] = "1"
        def chunks(k=k):
            step = (B + k - 1) // k
            for b in range(0, B, step): dev[b:b + step].copy_(host[b:b + step], non_blocking=True)
        t = run(chunks); print(f"{name} padded in {k} row blocks       {t*1e3:7.3f} ms {B*T*4/t/1e9:6.1f} GB/s (all bytes)")

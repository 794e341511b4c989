"""Probe: ragged H2D upload strategies (one stream, several streams, one contiguous copy)."""
import torch, time
torch.manual_seed(0)
B, T = 64, 320000
lens = torch.randint(32000, 320001, (B,)).sort(descending=True).values.tolist()
host = torch.randn(B, T).pin_memory()
dev = torch.empty(B, T, device="cuda")
tot = sum(lens) * 4
def run(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n
def one_stream():
    for b, n in enumerate(lens): dev[b, :n].copy_(host[b, :n], non_blocking=True)
streams = [torch.cuda.Stream() for _ in range(4)]
def multi(k):
    def f():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for s in streams[:k]: s.wait_event(ev)
        for b, n in enumerate(lens):
            with torch.cuda.stream(streams[b % k]): dev[b, :n].copy_(host[b, :n], non_blocking=True)
        for s in streams[:k]:
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
    return f
flat_h = torch.randn(sum(lens)).pin_memory(); flat_d = torch.empty(sum(lens), device="cuda")
def contiguous(): flat_d.copy_(flat_h, non_blocking=True)
def padded(): dev.copy_(host, non_blocking=True)
for name, fn in [("one stream, 64 ragged copies", one_stream), ("2 streams", multi(2)), ("4 streams", multi(4)),
                 ("one contiguous copy (same bytes)", contiguous)]:
    t = run(fn); print(f"{name:36s} {t*1e3:7.3f} ms  {tot/t/1e9:6.1f} GB/s")
t = run(padded); print(f"{'padded (B,T) copy':36s} {t*1e3:7.3f} ms  {B*T*4/t/1e9:6.1f} GB/s (all bytes)")

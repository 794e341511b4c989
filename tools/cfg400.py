"""Development: the n_fft = win_length = 400 recipe configuration (plan 11) on the C1 / C2 shapes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
dev = torch.device("cuda", 0)
def timeit(fn, n=200):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name in ("c1", "c2"):
    w = bench.WORKLOADS[name]
    lens = bench.workload_lengths(w)
    T = max(lens)
    ring = max(2, int(np.ceil(3 * bench.L2_BYTES / (4 * len(lens) * T * 1.33))) + 1)
    batches = [bench.make_batch(w, lens, 7 + i, dev)[..., 0] for i in range(ring)]
    wl = torch.tensor(lens, dtype=torch.int64, device=dev)
    host = np.asarray(lens, dtype=np.int64)
    m = Speech2MelSpec(n_mels=80, n_fft=400, win_length=400, hop_length=160, sr=16000, preemphasis=0.97 if name == "c2" else None).to(dev)
    plan = m.speech2linear._plan(dev, preemph=True, mel=m.linear2mel.mel_conf())
    for stats, skip in ((False, False), (True, True)):
        us = timeit(lambda i: plan.forward(batches[i % ring], wl, want_stats=stats, skip_pad=skip, len_host=host))
        print(f"n_fft 400 {name} stats={int(stats)} skip_pad={int(skip)}  {us:7.2f} us")
    del batches

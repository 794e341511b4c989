"""Development: where the C2 kernel's extra time per pair (against C1) comes from — ragged lengths, pre-emphasis,
fixed-point sums / skipped padded rows."""
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
w2 = bench.WORKLOADS["c2"]
lens_r = bench.workload_lengths(w2)
tot = sum(lens_r)
lens_f = [tot // 64] * 64  # the same audio in equal lengths
for lname, lens in (("ragged", lens_r), ("equal ", lens_f)):
    T = max(lens)
    ring = max(2, int(np.ceil(3 * bench.L2_BYTES / (4 * len(lens) * T * 1.33))) + 1)
    batches = [bench.make_batch(w2, lens, 7 + i, dev)[..., 0] for i in range(ring)]
    wl = torch.tensor(lens, dtype=torch.int64, device=dev)
    host = np.asarray(lens, dtype=np.int64)
    pairs = sum((1 + l // 160 + 1) // 2 for l in lens)
    for pre in (None, 0.97):
        m = Speech2MelSpec(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=pre).to(dev)
        plan = m.speech2linear._plan(dev, preemph=True, mel=m.linear2mel.mel_conf())
        for stats, skip in ((False, False), (False, True), (True, True)):
            us = timeit(lambda i: plan.forward(batches[i % ring], wl, want_stats=stats, skip_pad=skip, len_host=host))
            print(f"{lname} pre={pre} stats={int(stats)} skip_pad={int(skip)}  {us:7.2f} us  {us * 1e3 / pairs * 148:6.1f} ns/pair/SM ({pairs} real pairs)")
    del batches

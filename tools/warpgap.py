"""Development: the recipes' own SpecAugment block (time warp on by default, 4 frequency / 5 time masks) on the C2 shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.module.acoustic_frontend import AcousticFrontend
dev = torch.device("cuda", 0)
def timeit(fn, n=200):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
w = bench.WORKLOADS["c2"]
lens = bench.workload_lengths(w)
ring = 5
batches = [bench.make_batch(w, lens, 7 + i, dev) for i in range(ring)]
wl = torch.tensor(lens, dtype=torch.int64)
recipe = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
for name, norm, aug, ep in (("utterance norm + masks, no warp (C2)", dict(norm_type="utterance"), w["specaug"], None),
                            ("utterance norm + recipe SpecAugment (warp)", dict(norm_type="utterance"), recipe, None),
                            ("global norm, averaging (epoch 1) + recipe SpecAugment", dict(norm_type="global"), recipe, 1),
                            ("global norm, frozen (epoch 5) + recipe SpecAugment", dict(norm_type="global"), recipe, 5),
                            ("global norm, frozen (epoch 5) + recipe masks, no warp", dict(norm_type="global"), dict(recipe, time_warp=False), 5)):
    chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=norm, specaug=aug).to(dev)
    chain.train()
    for i in range(3): chain(batches[i], wl, epoch=1)
    us = timeit(lambda i: chain(batches[i % ring], wl, epoch=ep))
    import time
    t0 = time.perf_counter()
    for i in range(300): chain(batches[i % ring], wl, epoch=ep)
    host = (time.perf_counter() - t0) / 300 * 1e6
    torch.cuda.synchronize()
    print(f"{name:60s} {us:7.2f} us per step (host enqueue {host:5.1f} us)")

# the library call alone (fixed draws): what the device needs
chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=dict(norm_type="global", max_epoch_num=1), specaug=recipe).to(dev)
chain.train()
for i in range(3): chain(batches[i], wl, epoch=1)
fe = chain.frontend
plan = fe.speech2linear._plan(dev, preemph=True, mel=fe.linear2mel.mel_conf())
wl_dev = wl.to(dev)
host = wl.numpy()
flen_min = int((host.min() // 160) + 1)
warp, masks = chain.specaug.draw_device(len(lens), 80, None, dev, minlen=flen_min)
mean, std = chain.normalize.global_mean, chain.normalize.global_std
print("warp", warp)
for name, kw in (("chain mode 1 + masks", dict(norm_mode=1, mean=mean, std=std, masks=masks)),
                 ("chain mode 1 + masks + warp", dict(norm_mode=1, mean=mean, std=std, masks=masks, warp=warp)),
                 ("chain mode 2 + masks", dict(norm_mode=2, masks=masks)),
                 ("chain mode 2 + masks + warp", dict(norm_mode=2, masks=masks, warp=warp))):
    us = timeit(lambda i: plan.chain(batches[i % ring], wl_dev, len_host=host, **kw))
    print(f"{name:52s} {us:7.2f} us per call")

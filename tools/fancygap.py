"""Development: the frontend kernel with its in-kernel normalise + mask epilogue (statistics known up front: evaluation,
or global statistics frozen after max_epoch_num) against the plain launch, C2 shape."""
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
T = max(lens)
ring = 5
batches = [bench.make_batch(w, lens, 7 + i, dev) for i in range(ring)]
wl = torch.tensor(lens, dtype=torch.int64)
for name, norm, aug, train in (("no norm, no masks (eval)", None, None, False),
                               ("no norm + masks (train)", None, w["specaug"], True),
                               ("global norm frozen, eval", dict(norm_type="global"), None, False),
                               ("global norm frozen + masks (train, epoch > max)", dict(norm_type="global", max_epoch_num=1), w["specaug"], True),
                               ("utterance norm + masks (C2)", dict(norm_type="utterance"), w["specaug"], True)):
    chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=norm, specaug=aug).to(dev)
    chain.train()
    if norm and norm["norm_type"] == "global":
        for i in range(3): chain(batches[i], wl, epoch=1)   # accumulate statistics
    chain.train(train)
    ep = 5 if train else None
    us = timeit(lambda i: chain(batches[i % ring], wl, epoch=ep))
    print(f"{name:52s} {us:7.2f} us per step")

# the launch alone: plan.chain with fixed arguments (no draws, no module logic)
from speechain_b200.host import ops
chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=dict(norm_type="global", max_epoch_num=1), specaug=w["specaug"]).to(dev)
chain.train()
for i in range(3): chain(batches[i], wl, epoch=1)
fe = chain.frontend
plan = fe.speech2linear._plan(dev, preemph=True, mel=fe.linear2mel.mel_conf())
wl_dev = wl.to(dev)
host = wl.numpy()
flen_min = int((host.min() // 160) + 1)
_, masks = chain.specaug.draw_device(len(lens), 80, None, dev, minlen=flen_min)
mean, std = chain.normalize.global_mean, chain.normalize.global_std
for name, kw in (("chain mode 0 (plain)", dict(norm_mode=0)),
                 ("chain mode 0 + masks", dict(norm_mode=0, masks=masks)),
                 ("chain mode 1 (mean/std)", dict(norm_mode=1, mean=mean, std=std)),
                 ("chain mode 1 + masks", dict(norm_mode=1, mean=mean, std=std, masks=masks))):
    us = timeit(lambda i: plan.chain(batches[i % ring], wl_dev, len_host=host, **kw))
    print(f"{name:52s} {us:7.2f} us per launch")

# which part of the masks costs: frequency / time masks alone, with and without the rows past each length
import numpy as np
from speechain_b200.host.ops import DeviceMasks
B = len(lens)
rng = np.random.default_rng(0)
fpos, flen_ = rng.integers(0, 50, (B, 2)), rng.integers(0, 30, (B, 2))
tpos, tlen_ = rng.integers(0, 150, (B, 2)), rng.integers(0, 10, (B, 2))
for name, mk in (("freq masks only", DeviceMasks(dev, fpos, flen_, None, None)),
                 ("time masks only", DeviceMasks(dev, None, None, tpos, tlen_)),
                 ("both", DeviceMasks(dev, fpos, flen_, tpos, tlen_))):
    for skip in (False, True):
        us = timeit(lambda i: plan.forward(batches[i % ring][..., 0], wl_dev, masks=mk, skip_pad=skip, len_host=host))
        print(f"forward + {name:16s} skip_pad={int(skip)}                     {us:7.2f} us per launch")
for skip in (False, True):
    us = timeit(lambda i: plan.forward(batches[i % ring][..., 0], wl_dev, skip_pad=skip, len_host=host))
    print(f"forward, no masks        skip_pad={int(skip)}                     {us:7.2f} us per launch")
    us = timeit(lambda i: plan.forward(batches[i % ring][..., 0], wl_dev, norm_mean=mean, norm_std=std, skip_pad=skip, len_host=host))
    print(f"forward, mean/std        skip_pad={int(skip)}                     {us:7.2f} us per launch")

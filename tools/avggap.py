"""Development: the C2 batch through the chain while the global statistics are still being averaged (epoch <=
max_epoch_num: frontend launch with sums -> finalise -> running update -> normalise + masks), device time per step, host
enqueue time per step and a cProfile of the host side."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.module.acoustic_frontend import AcousticFrontend
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["c2"]
lens = bench.workload_lengths(w)
ring = 5
batches = [bench.make_batch(w, lens, 7 + i, dev) for i in range(ring)]
wl = torch.tensor(lens, dtype=torch.int64)
recipe = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
for name, aug, kind in (("averaging epoch, no SpecAugment", None, "global"), ("averaging epoch + bench masks", w["specaug"], "global"),
                        ("averaging epoch + recipe SpecAugment (time warp)", recipe, "global"),
                        ("batch statistics (training) + bench masks", w["specaug"], "batch")):
    chain = AcousticFrontend(frontend=dict(conf=w["frontend"]), normalize=dict(norm_type=kind), specaug=aug).to(dev)
    chain.train()
    for i in range(5): chain(batches[i % ring], wl, epoch=1)
    torch.cuda.synchronize()
    n = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter()
    for i in range(n): chain(batches[i % ring], wl, epoch=1)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"{name:52s} device {e0.elapsed_time(e1) / n * 1e3:7.2f} us per step, host enqueue {(t1 - t0) / n * 1e6:7.2f} us")
    if aug is None or "--all" in sys.argv:
        pr = cProfile.Profile(); pr.enable()
        for i in range(n): chain(batches[i % ring], wl, epoch=1)
        pr.disable(); torch.cuda.synchronize()
        st = pstats.Stats(pr); st.sort_stats("tottime"); st.print_stats(22)

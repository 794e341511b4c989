"""Development: the C2 batch through the three modules one after the other, as the reference's ASREncoder.forward calls
them (encoder/asr.py:103-119) when only the module files are swapped — against AcousticFrontend's fused schedules."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
from speechain_b200.module.norm.feat_norm import FeatureNormalization
from speechain_b200.module.augment.specaug import SpecAugment
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["c2"]
lens = bench.workload_lengths(w)
ring = 5
batches = [bench.make_batch(w, lens, 7 + i, dev) for i in range(ring)]
wl = torch.tensor(lens, dtype=torch.int64, device=dev)
recipe = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
for name, nconf, aug, epoch in (("utterance norm + bench masks", dict(norm_type="utterance"), w["specaug"], 1),
                                ("global (averaging) + recipe SpecAugment", dict(norm_type="global"), recipe, 1),
                                ("global (frozen) + recipe SpecAugment", dict(norm_type="global", max_epoch_num=0), recipe, 5)):
    fe = Speech2MelSpec(**w["frontend"]).to(dev)
    norm = FeatureNormalization(input_size=80, **nconf).to(dev)
    sa = SpecAugment(input_size=80, feat_norm=True, **aug).to(dev)
    for m in (fe, norm, sa): m.train()
    def step(i):
        y, yl = fe(batches[i % ring], wl)
        y, yl = norm(y, yl, epoch=epoch if epoch != 5 or i > 2 else 0)
        return sa(y, yl)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    n = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); t0 = time.perf_counter()
    for i in range(5, n + 5): step(i)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"modules one by one: {name:44s} device {e0.elapsed_time(e1) / n * 1e3:7.2f} us per step, host enqueue {(t1 - t0) / n * 1e6:7.2f} us")
    if "--prof" in sys.argv:
        pr = cProfile.Profile(); pr.enable()
        for i in range(5, n + 5): step(i)
        pr.disable(); torch.cuda.synchronize()
        st = pstats.Stats(pr); st.sort_stats("tottime"); st.print_stats(18)

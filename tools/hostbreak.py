"""Per-function host cost of one AcousticFrontend step (nested wall-clock timers, no profiler)."""
import sys, time, collections, torch
sys.path.insert(0, '.')
import speechain_b200.module.acoustic_frontend as af
from speechain_b200.host import ops
from speechain_b200.module.augment.specaug import SpecAugment
from speechain_b200.module.frontend.speech2linear import Speech2LinearSpec

acc = collections.defaultdict(float)
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] += time.perf_counter() - t
    return w

dev = 'cuda:0'
conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
chain = af.AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"),
                            specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)).to(dev)
chain.train()
B = 64
wav = torch.randn(B, 8000, 1, device=dev)
wl = torch.full((B,), 8000, dtype=torch.int64, device=dev)
for _ in range(50): chain(wav, wl)
torch.cuda.synchronize()
N = 1000
t0 = time.perf_counter()
for _ in range(N): chain(wav, wl)
base = (time.perf_counter() - t0) / N * 1e6
torch.cuda.synchronize()
af.squeeze_speech = timed("squeeze_speech", af.squeeze_speech)
af.lengths_pair = timed("lengths_pair", af.lengths_pair)
Speech2LinearSpec._feat_len_host = timed("_feat_len_host", Speech2LinearSpec._feat_len_host)
Speech2LinearSpec._feat_len_of = timed("_feat_len_of", Speech2LinearSpec._feat_len_of)
Speech2LinearSpec._plan = timed("_plan", Speech2LinearSpec._plan)
SpecAugment.draw_device = timed("draw_device", SpecAugment.draw_device)
SpecAugment.draw = timed("  draw", SpecAugment.draw)
ops.DeviceMasks.__init__ = timed("  DeviceMasks()", ops.DeviceMasks.__init__)
ops.DeviceMasks.struct = timed("  masks.struct", ops.DeviceMasks.struct)
ops.FrontendPlan.chain = timed("plan.chain", ops.FrontendPlan.chain)
ops.FrontendPlan.num_frames = timed("  num_frames", ops.FrontendPlan.num_frames)
ops.HostLengths.remember = classmethod(timed("remember", ops.HostLengths.remember.__func__))
plan = chain.frontend.speech2linear._plan(torch.device(dev), preemph=True, mel=chain.frontend.linear2mel.mel_conf())
raw = plan._lib.sfe_chain_forward
class LibProxy:
    def __init__(self, lib): self._l = lib; self.sfe_chain_forward = timed("  ctypes sfe_chain_forward", raw)
    def __getattr__(self, k): return getattr(self._l, k)
plan._lib = LibProxy(plan._lib)
_empty = torch.empty
torch.empty = timed("  torch.empty (x2)", _empty)
fwd = timed("forward", chain.forward)
for _ in range(N): fwd(wav, wl)
torch.cuda.synchronize()
print(f"untimed host per step {base:.1f} us")
for k, v in acc.items():
    print(f"{k:32s} {v / N * 1e6:7.2f} us")

import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import frontend_np as O
from speechain_b200.module.acoustic_frontend import AcousticFrontend
from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
DEV='cuda:0'
g = torch.Generator().manual_seed(1236)
lens = sorted(torch.randint(32000, 320001, (64,), generator=g).tolist(), reverse=True)
gg = torch.Generator().manual_seed(1236)
wav = 0.1 * torch.randn(len(lens), max(lens), generator=gg)
for i, n in enumerate(lens): wav[i, n:] = 0
wav[0, 1000:9000] = 0
conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"),
                         specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)).to(DEV)
chain.train()
torch.manual_seed(1234)
y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
flen = O.feat_lengths(np.array(lens), 512, 160, True)
d = O.specaug_draw(O.TorchCpuRandint(1234), 64, 80, flen, O.SpecAugCfg(time_warp=False))
mask = O.specaug_mask(64, y.shape[1], 80, d)
got = y.cpu().numpy()
z = (got == 0.0)
bad = np.argwhere(z != mask)
print("mismatches", len(bad), "shape", got.shape)
print(bad[:10])
for b_, t_, d_ in bad[:5]:
    print(b_, t_, d_, "got", got[b_, t_, d_], "mask", mask[b_, t_, d_], "flen", flen[b_])
print("nan count", np.isnan(got).sum())
# raw frontend check
m = Speech2MelSpec(**conf).to(DEV)
f, fl = m(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
f = f.cpu().numpy()
print("raw finite", np.isfinite(f).all(), "min", f.min(), "max", f.max())
ub = np.unique(bad[:,0]); print("bad utts", ub[:20])

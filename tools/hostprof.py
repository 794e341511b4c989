import sys, time, cProfile, pstats, torch
sys.path.insert(0, '.')
from speechain_b200.module.acoustic_frontend import AcousticFrontend
dev = 'cuda:0'
conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"),
                         specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)).to(dev)
chain.train()
B = 64
wav = torch.randn(B, 8000, 1, device=dev)
wl = torch.full((B,), 8000, dtype=torch.int64, device=dev)
for _ in range(50): chain(wav, wl)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 500
for _ in range(N): chain(wav, wl)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host per step {1e6*(t1-t0)/N:.1f} us; incl drain {1e6*(t2-t0)/N:.1f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): chain(wav, wl)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)

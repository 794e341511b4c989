"""Generate the golden fixtures in this directory from the *reference itself*.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the five leaf modules of bagustris/speechain from /root/reference (they need
only torch + torchaudio), runs them on small seeded inputs on CPU, and stores inputs and
outputs as ``*.npz``.  The reference has no tests for this path, so these
files are the known-answer set that pins ``oracle/`` and the CUDA path.  The script also
checks that ``oracle/ref_port_torch.py`` reproduces the reference bit for bit.

``FeatureNormalization`` in ``global`` training mode calls ``.cuda()`` at feat_norm.py:475
and cannot run on CPU; ``_CpuFeatNorm`` overrides *only* ``register_mean_std_batch`` with
the same arithmetic minus that call.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from speechain.module.augment.specaug import SpecAugment  # noqa: E402
from speechain.module.frontend.linear2mel import LinearSpec2MelSpec  # noqa: E402
from speechain.module.frontend.speech2linear import Speech2LinearSpec  # noqa: E402
from speechain.module.frontend.speech2mel import Speech2MelSpec  # noqa: E402
from speechain.module.norm.feat_norm import FeatureNormalization  # noqa: E402

from oracle import ref_port_torch as port  # noqa: E402


class _CpuFeatNorm(FeatureNormalization):
    def register_mean_std_batch(self, curr_aver_mean, curr_aver_std, prefix, epoch):
        if epoch is None or not hasattr(self, f"{prefix}_batch"):
            self.register_buffer(f"{prefix}_batch", torch.LongTensor([1]))
        elif epoch <= self.max_epoch_num:
            self.register_buffer(f"{prefix}_batch", self.get_buffer(f"{prefix}_batch") + 1)
        if epoch is None or epoch <= self.max_epoch_num:
            w = 1 / self.get_buffer(f"{prefix}_batch")
            if self.mean_norm:
                if not hasattr(self, f"{prefix}_mean"):
                    self.register_buffer(f"{prefix}_mean", curr_aver_mean)
                else:
                    self.register_buffer(f"{prefix}_mean", w * curr_aver_mean + (1 - w) * self.get_buffer(f"{prefix}_mean"))
            if self.std_norm:
                if not hasattr(self, f"{prefix}_std"):
                    self.register_buffer(f"{prefix}_std", curr_aver_std)
                else:
                    self.register_buffer(f"{prefix}_std", w * curr_aver_std + (1 - w) * self.get_buffer(f"{prefix}_std"))


def make_wave(seed, lens, zero_seg=None, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    T = max(lens)
    wav = scale * torch.randn(len(lens), T, 1, generator=g)
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    if zero_seg is not None:
        wav[0, zero_seg[0]:zero_seg[1]] = 0
    return wav, torch.tensor(lens, dtype=torch.int64)


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]  # optional: write only fixtures whose name starts with one of these


def save(name, **arrs):
    if ONLY and not any(name.startswith(o) for o in ONLY):
        return
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"  wrote {name}.npz  ({sum(a.nbytes for a in out.values())/1e3:.0f} kB raw)")


def bitexact(a, b, what):
    same = torch.equal(a, b)
    print(f"  port == reference ({what}): {same}")
    assert same, what


def frontend_case(name, lens, seed, conf, zero_seg=None, port_check=True):
    wav, wl = make_wave(seed, lens, zero_seg)
    m = Speech2MelSpec(**conf)
    out = m(wav.clone(), wl.clone())
    feat, fl = out[0], out[1]
    extra = {}
    if len(out) == 4:
        extra = dict(energy=out[2])
    if port_check:
        sr = conf.get("sr", 16000)
        hop = int(conf["hop_length"] * sr) if isinstance(conf["hop_length"], float) else conf["hop_length"]
        win = int(conf["win_length"] * sr) if isinstance(conf["win_length"], float) else conf["win_length"]
        n_fft = conf.get("n_fft") or win
        spec, pfl, pen = port.speech2linear(wav.clone(), wl.clone(), n_fft=n_fft, hop=hop, win=win,
                                            window=conf.get("window", "hann"), center=conf.get("center", True),
                                            normalized=conf.get("normalized", False),
                                            preemphasis=conf.get("preemphasis"), mag_spec=conf.get("mag_spec", False),
                                            return_energy=conf.get("return_energy", False))
        W = port.mel_weight(sr, n_fft, conf["n_mels"], conf.get("fmin", 0.0), conf.get("fmax"),
                            conf.get("mel_scale", "slaney"), conf.get("mel_norm", True))
        bitexact(W, m.linear2mel.mel_fbanks.weight, name + " mel weight")
        pf = port.linear2mel(spec, W, clamp=conf.get("clamp", 1e-10), logging=conf.get("logging", True),
                             log_base=conf.get("log_base", 10.0))
        if conf.get("delta_order") is None:
            bitexact(pf, feat, name + " feat")
        bitexact(pfl, fl, name + " feat_len")
        if pen is not None:
            bitexact(pen, out[2], name + " energy")
    save(name, wav=wav[..., 0], wav_len=wl, feat=feat, feat_len=fl, mel_weight=m.linear2mel.mel_fbanks.weight, **extra)
    return wav, wl, feat, fl


def main():
    torch.set_num_threads(1)
    print("frontend cases")
    C1 = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000)
    frontend_case("fe_c1", [6000, 6000, 4321, 1600], 1235, C1, zero_seg=(1000, 3000))
    frontend_case("fe_c2_512", [6400, 5000, 3211, 800], 1236, dict(C1, preemphasis=0.97), zero_seg=(500, 2500))
    frontend_case("fe_c2_400", [6400, 5000, 3211, 800], 1236,
                  dict(n_mels=80, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97), zero_seg=(500, 2500))
    C3 = dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000,
              log_base=None, clamp=1e-5, return_energy=True)
    frontend_case("fe_c3", [9000, 9000, 5555], 1237, C3, zero_seg=(2000, 6000))
    frontend_case("fe_htk_nonorm", [5000, 3000], 1240,
                  dict(n_mels=40, n_fft=512, win_length=400, hop_length=160, mel_scale="htk", mel_norm=False,
                       window="hamming", normalized=True))
    frontend_case("fe_nocenter", [5000, 3100], 1241,
                  dict(n_mels=80, n_fft=512, win_length=400, hop_length=160, center=False, logging=False))
    frontend_case("fe_nfft256", [3000, 2000, 999], 1242, dict(n_mels=40, n_fft=256, win_length=200, hop_length=80, sr=8000))
    frontend_case("fe_nfft2048", [12000, 7000], 1243, dict(n_mels=128, n_fft=2048, win_length=1200, hop_length=300, sr=24000))
    frontend_case("fe_delta2", [4000, 2500], 1244, dict(C1, delta_order=2), port_check=True)
    # waveform pre-normalisation (no shipped recipe uses it): min_max for a batch, mean_std for one utterance
    frontend_case("fe_prenorm_minmax", [4000, 2600, 1700], 1246, dict(C1, preemphasis=0.97, pre_stft_norm="min_max"),
                  port_check=False)
    frontend_case("fe_prenorm_meanstd", [4000], 1247, dict(C1, pre_stft_norm="mean_std"), port_check=False)
    # odd T_max / odd lengths: row pitch not a multiple of 4 samples
    frontend_case("fe_oddpitch", [5003, 5003, 4999, 1234, 161], 1245, dict(C1, preemphasis=0.97))

    print("linear-spectrogram cases (Speech2LinearSpec standalone)")
    wav, wl = make_wave(1250, [7000, 4100, 2000])
    for nm, conf in [("lin_mag_log", dict(sr=22050, win_length=1024, hop_length=256, mag_spec=True, logging=True, clamp=1e-5)),
                     ("lin_pow", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, return_energy=True)),
                     ("lin_pow_log10", dict(sr=16000, win_length=400, hop_length=160, preemphasis=0.97, logging=True, log_base=10.0)),
                     # all n_fft bins (torch.stft(onesided=False)); the energy then sums the mirrored half too
                     ("lin_twosided", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False, return_energy=True)),
                     ("lin_twosided_mag_log", dict(sr=16000, win_length=400, hop_length=160, onesided=False, mag_spec=True,
                                                   logging=True, clamp=1e-5, return_energy=True))]:
        m = Speech2LinearSpec(**conf)
        out = m(wav.clone(), wl.clone())
        extra = dict(energy=out[2]) if len(out) == 4 else {}
        save(nm, wav=wav[..., 0], wav_len=wl, spec=out[0], feat_len=out[1], **extra)

    print("linear -> mel standalone")
    g = torch.Generator().manual_seed(1260)
    spec = torch.rand(2, 30, 257, generator=g) * 3
    m = LinearSpec2MelSpec(n_fft=512, n_mels=80)
    mel, _ = m(spec, torch.tensor([30, 20]))
    save("l2m", spec=spec, mel=mel, mel_weight=m.mel_fbanks.weight)

    print("mel -> linear recovery (LinearSpec2MelSpec.recover, linear2mel.py:191-224)")
    flen = torch.tensor([30, 20])
    rec = m.recover(mel.clone(), flen.clone())
    save("m2l_log10", mel=mel, feat_len=flen, lin=rec, mel_weight=m.mel_fbanks.weight)
    g = torch.Generator().manual_seed(1261)
    spec2 = torch.rand(3, 24, 513, generator=g) ** 3 * 5
    m2 = LinearSpec2MelSpec(n_fft=1024, n_mels=80, sr=22050, fmin=0, fmax=8000, clamp=1e-5, log_base=None)
    flen2 = torch.tensor([24, 24, 11])
    mel2, _ = m2(spec2, flen2.clone())
    rec2 = m2.recover(mel2.clone(), flen2.clone())
    save("m2l_ln_tts", mel=mel2, feat_len=flen2, lin=rec2, mel_weight=m2.mel_fbanks.weight,
         conf=np.array([1024, 80, 22050, 8000]))
    m3 = LinearSpec2MelSpec(n_fft=512, n_mels=80, logging=False)
    mel3, _ = m3(spec, flen.clone())
    rec3 = m3.recover(mel3.clone(), flen.clone())
    save("m2l_nolog", mel=mel3, feat_len=flen, lin=rec3, mel_weight=m3.mel_fbanks.weight)

    print("Griffin-Lim recovery (Speech2LinearSpec.recover / Speech2MelSpec.recover)")
    import math

    def gl_case(name, conf, lens, seed, mel=False):
        sr = conf.get("sr", 16000)
        T = max(lens)
        t = torch.arange(T) / sr
        gg = torch.Generator().manual_seed(seed)
        base = 0.3 * torch.sin(2 * math.pi * 220 * t) + 0.2 * torch.sin(2 * math.pi * 1333 * t) + 0.02 * torch.randn(T, generator=gg)
        wav = torch.stack([base * (0.9 ** i) if i % 2 == 0 else base.flip(0) * 0.7 for i in range(len(lens))])[..., None]
        wl = torch.tensor(lens)
        for i, n in enumerate(lens):
            wav[i, n:] = 0
        cls = Speech2MelSpec if mel else Speech2LinearSpec
        feat, fl = cls(**conf)(wav.clone(), wl.clone())[:2]
        fresh = cls(**conf)  # recover() only works on a module whose forward never ran (it re-reads stft_config["window"])
        K = (conf.get("n_fft") or conf["win_length"]) // 2 + 1
        torch.manual_seed(seed)
        angles = torch.rand(len(lens), K, feat.shape[1], dtype=torch.cfloat)
        torch.manual_seed(seed)
        rec, rl = fresh.recover(feat.clone(), fl.clone())
        save(name, feat=feat, feat_len=fl, angles=angles.transpose(1, 2).contiguous(), wav=rec, wav_len=rl, seed=np.array(seed))

    gl_case("gl_lin_512", dict(hop_length=160, win_length=400, n_fft=512, sr=16000), [8000, 6400], 11)
    gl_case("gl_lin_pre_400", dict(hop_length=160, win_length=400, sr=16000, preemphasis=0.97, mag_spec=True), [6400, 6400, 3200], 12)
    gl_case("gl_mel_tts_1024", dict(n_mels=80, hop_length=256, win_length=1024, sr=22050, fmin=0, fmax=8000, clamp=1e-5,
                                     log_base=None, preemphasis=0.97), [11025, 8000], 13, mel=True)

    print("utterance norm + SpecAugment (C2 chain)")
    wav, wl, feat, fl = frontend_case("chain_c2_fe", [9600, 8000, 6400, 4800, 3300, 3300], 1270, dict(C1, preemphasis=0.97))
    un = FeatureNormalization(input_size=80, norm_type="utterance")
    nfeat, _ = un(feat.clone(), fl.clone())
    bitexact(port.norm_utterance(feat.clone(), fl), nfeat, "utterance norm")
    sa = SpecAugment(input_size=80, time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                     time_mask_num=2, feat_norm=True)
    sa.train()
    torch.manual_seed(1234)
    afeat, _ = sa(nfeat.clone(), fl.clone())
    torch.manual_seed(1234)
    bitexact(port.specaug(nfeat.clone(), fl), afeat, "specaug")
    save("chain_c2", feat=feat, feat_len=fl, normed=nfeat, augmented=afeat, seed=1234)

    print("SpecAugment variants")
    g = torch.Generator().manual_seed(1280)
    x = torch.randn(5, 60, 80, generator=g)
    xl = torch.tensor([60, 55, 41, 33, 33])
    for nm, conf, seed in [
        ("sa_recipe", dict(time_warp=False, freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5), 11),
        ("sa_intwidth", dict(time_warp=False, freq_mask_width=[3, 9], freq_mask_num=3, time_mask_width=20, time_mask_num=4), 12),
        ("sa_nonorm", dict(time_warp=False, feat_norm=False), 13),
        ("sa_freq_only", dict(time_warp=False, time_mask=False), 14),
        ("sa_time_only", dict(time_warp=False, freq_mask=False, time_mask_width=[0.02, 0.2]), 15),
        ("sa_warp", dict(time_warp=True), 16),
        ("sa_warp_only", dict(time_warp=True, freq_mask=False, time_mask=False, time_warp_window=8), 17),
        ("sa_warp_bilinear", dict(time_warp=True, time_warp_mode="bilinear"), 19),
    ]:
        sa = SpecAugment(input_size=80, **conf)
        sa.train()
        torch.manual_seed(seed)
        y, _ = sa(x.clone(), xl.clone())
        save(nm, x=x, x_len=xl, y=y, seed=seed)
    # too short for warping: minlen <= 2*window+1
    xs = torch.randn(3, 14, 80, generator=g)
    xsl = torch.tensor([14, 12, 11])
    sa = SpecAugment(input_size=80)
    sa.train()
    torch.manual_seed(18)
    y, _ = sa(xs.clone(), xsl.clone())
    save("sa_short", x=xs, x_len=xsl, y=y, seed=18)

    print("global / batch normalisation")
    g = torch.Generator().manual_seed(1290)
    batches = []
    for _ in range(6):
        f = torch.randn(4, 50, 80, generator=g) * 2 - 5
        l = torch.tensor([50, 44, 30, 17])
        batches.append((f, l))
    gn = _CpuFeatNorm(input_size=80, norm_type="global")
    gn.train()
    epochs = [1, 1, 2, 5, 6, None]
    rec = {}
    for i, ((f, l), ep) in enumerate(zip(batches, epochs)):
        y, _ = gn(f.clone(), l.clone(), epoch=ep)
        rec[f"x{i}"], rec[f"len{i}"], rec[f"y{i}"] = f, l, y
        rec[f"mean{i}"], rec[f"std{i}"], rec[f"batch{i}"] = gn.global_mean.clone(), gn.global_std.clone(), gn.global_batch.clone()
    gn.eval()
    y, _ = gn(batches[0][0].clone(), batches[0][1].clone())
    rec["y_eval"] = y
    rec["epochs"] = np.array([-1 if e is None else e for e in epochs])
    rec["state_keys"] = np.array(sorted(gn.state_dict().keys()))
    save("gnorm_seq", **rec)
    # port check of the single-process global path
    st = {}
    for (f, l), ep in zip(batches, epochs):
        py = port.norm_global_train(f.clone(), l, st, ep)
    gn2 = _CpuFeatNorm(input_size=80, norm_type="global")
    gn2.train()
    for (f, l), ep in zip(batches, epochs):
        ry, _ = gn2(f.clone(), l.clone(), epoch=ep)
    bitexact(py, ry, "global norm sequence")

    bn = _CpuFeatNorm(input_size=80, norm_type="batch")
    bn.train()
    yt, _ = bn(batches[1][0].clone(), batches[1][1].clone())
    bn.eval()
    ye, _ = bn(batches[1][0].clone(), batches[1][1].clone())
    save("bnorm", x=batches[1][0], x_len=batches[1][1], y_train=yt, y_eval=ye)

    # mean-only / std-only and 1-D (B, L) inputs (pitch / energy normalisers of FastSpeech2)
    for nm, conf in [("unorm_meanonly", dict(norm_type="utterance", std_norm=False)),
                     ("unorm_stdonly", dict(norm_type="utterance", mean_norm=False))]:
        m = FeatureNormalization(input_size=80, **conf)
        y, _ = m(batches[2][0].clone(), batches[2][1].clone())
        save(nm, x=batches[2][0], x_len=batches[2][1], y=y)
    e = torch.rand(4, 50, generator=g) * 3
    el = torch.tensor([50, 44, 30, 17])
    m = FeatureNormalization(input_size=1, norm_type="utterance")
    y, _ = m(e.clone(), el.clone())
    g1 = _CpuFeatNorm(input_size=1, norm_type="global")
    g1.train()
    yg, _ = g1(e.clone(), el.clone(), epoch=1)
    save("norm_1d", x=e, x_len=el, y_utt=y, y_global=yg)

    # group normalisation: tensor ids (speaker-wise) and a str id (domain-wise)
    grp = _CpuFeatNorm(input_size=80, norm_type="group")
    grp.train()
    ids = [torch.tensor([3, 7, 3, 9]), torch.tensor([7, 7, 3, 11])]
    rec = {}
    for i in range(2):
        y, _ = grp(batches[i][0].clone(), batches[i][1].clone(), group_ids=ids[i], epoch=1)
        rec[f"x{i}"], rec[f"len{i}"], rec[f"ids{i}"], rec[f"y{i}"] = batches[i][0], batches[i][1], ids[i], y
    for k, v in grp.state_dict().items():
        rec["buf_" + k] = v
    grp.eval()
    y, _ = grp(batches[2][0].clone(), batches[2][1].clone(), group_ids=torch.tensor([3, 5, 7, 9]))
    rec["x_eval"], rec["len_eval"], rec["ids_eval"], rec["y_eval"] = batches[2][0], batches[2][1], torch.tensor([3, 5, 7, 9]), y
    save("grpnorm_tensor", **rec)
    gs = _CpuFeatNorm(input_size=80, norm_type="group")
    gs.train()
    y0, _ = gs(batches[0][0].clone(), batches[0][1].clone(), group_ids="real", epoch=1)
    y1, _ = gs(batches[1][0].clone(), batches[1][1].clone(), group_ids="syn", epoch=1)
    save("grpnorm_str", x0=batches[0][0], len0=batches[0][1], y0=y0, x1=batches[1][0], len1=batches[1][1], y1=y1,
         state_keys=np.array(sorted(gs.state_dict().keys())))

    print("done")


if __name__ == "__main__":
    main()

"""Pin the CPU oracle (oracle/frontend_np.py) against the golden fixtures that
tests/golden/make_golden.py produced by running the reference modules themselves."""
import numpy as np
import pytest

from oracle import frontend_np as O
from sfe_testutil import assert_gl_close, assert_close, load_golden

FE_CASES = {
    "fe_c1": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010),
    "fe_c2_512": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97),
    "fe_c2_400": dict(n_mels=80, win_length=0.025, hop_length=0.010, preemphasis=0.97),
    "fe_c3": dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000,
                  log_base=None, clamp=1e-5, return_energy=True),
    "fe_htk_nonorm": dict(n_mels=40, n_fft=512, win_length=400, hop_length=160, mel_scale="htk", mel_norm=False,
                          window="hamming", normalized=True),
    "fe_nocenter": dict(n_mels=80, n_fft=512, win_length=400, hop_length=160, center=False, logging=False),
    "fe_nfft256": dict(n_mels=40, n_fft=256, win_length=200, hop_length=80, sr=8000),
    "fe_nfft2048": dict(n_mels=128, n_fft=2048, win_length=1200, hop_length=300, sr=24000),
    "fe_delta2": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, delta_order=2),
    "fe_oddpitch": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97),
    "chain_c2_fe": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97),
    "fe_prenorm_minmax": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97,
                              pre_stft_norm="min_max"),
    "fe_prenorm_meanstd": dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, pre_stft_norm="mean_std"),
}


@pytest.mark.parametrize("name", sorted(FE_CASES))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_speech2mel(name, dtype):
    g = load_golden(name)
    cfg = O.FrontendCfg(**FE_CASES[name])
    feat, flen, energy = O.speech2mel(g["wav"].astype(dtype), g["wav_len"], cfg)
    assert (flen == g["feat_len"]).all()
    assert_close(feat, g["feat"], what=name)
    if "energy" in g:
        assert_close(energy, g["energy"], what=name + " energy")
    # padded frames carry log(clamp)/ln(base), not zero
    if cfg.logging and cfg.delta_order is None:
        pad = np.log(cfg.clamp) / (np.log(cfg.log_base) if cfg.log_base else 1.0)
        for i, n in enumerate(flen):
            if n < feat.shape[1]:
                assert np.allclose(feat[i, n:], pad, rtol=1e-6)


@pytest.mark.parametrize("name", sorted(FE_CASES))
def test_mel_filterbank(name):
    g = load_golden(name)
    c = O.FrontendCfg(**FE_CASES[name])
    W = O.mel_filterbank(c.sr, c.n_fft, c.n_mels, c.fmin, c.fmax, c.mel_scale, c.mel_norm)
    ref = g["mel_weight"]
    assert W.shape == ref.shape
    # the reference builds the bank in fp32 (torch.linspace etc.): agree to fp32 rounding of Hz values
    assert np.abs(W - ref).max() <= 4e-6 * max(1.0, np.abs(ref).max())
    # banded: every row's support is contiguous
    for row in ref:
        nz = np.nonzero(row)[0]
        if nz.size:
            assert nz[-1] - nz[0] + 1 == nz.size


@pytest.mark.parametrize("name,cfg", [
    ("lin_mag_log", dict(sr=22050, win_length=1024, hop_length=256, mag_spec=True, lin_logging=True, lin_clamp=1e-5)),
    ("lin_pow", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, return_energy=True)),
    ("lin_pow_log10", dict(sr=16000, win_length=400, hop_length=160, preemphasis=0.97, lin_logging=True, lin_log_base=10.0)),
    ("lin_twosided", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False, return_energy=True)),
    ("lin_twosided_mag_log", dict(sr=16000, win_length=400, hop_length=160, onesided=False, mag_spec=True, lin_logging=True,
                                  lin_clamp=1e-5, return_energy=True)),
])
def test_speech2linear(name, cfg):
    g = load_golden(name)
    spec, flen, energy = O.speech2linear(g["wav"].astype(np.float64), g["wav_len"], O.FrontendCfg(**cfg))
    assert (flen == g["feat_len"]).all()
    if name == "lin_pow_log10":
        # log of *single* STFT bins: bins emptied by the pre-emphasis null sit at the fp32 FFT noise
        # floor of the reference itself, where the log is ill-conditioned (the fp32 reference is
        # 2.4e-3 away from fp64 there).  Compare in the power domain against the frame's scale.
        p, pref = 10.0 ** spec, 10.0 ** g["spec"].astype(np.float64)
        assert_close(p, pref, rtol=1e-4, atol=1e-6 * pref.max(), what=name)
    elif name == "lin_twosided_mag_log":
        assert spec.shape[-1] == 400
        p, pref = np.exp(spec), np.exp(g["spec"].astype(np.float64))
        assert_close(p, pref, rtol=1e-4, atol=1e-5 * pref.max(), what=name)
    else:
        # power spectra of 0.1*randn reach ~1e2: compare with the north-star tolerance
        assert_close(spec, g["spec"], what=name)
    if "energy" in g:
        assert_close(energy, g["energy"], what=name + " energy")


def test_linear2mel_standalone():
    g = load_golden("l2m")
    mel = O.linear2mel(g["spec"].astype(np.float64), g["mel_weight"].astype(np.float64))
    assert_close(mel, g["mel"], what="l2m")


M2L_CASES = {
    "m2l_log10": dict(logging=True, log_base=10.0),
    "m2l_ln_tts": dict(logging=True, log_base=None),
    "m2l_nolog": dict(logging=False, log_base=10.0),
}


@pytest.mark.parametrize("name", sorted(M2L_CASES))
def test_mel2linear_recover(name):
    """mel -> linear least squares (linear2mel.py:191-224).  The solution cancels heavily inside a frame, so
    the comparison is relative to the frame's largest bin (fp32 lstsq noise of the reference ~5e-7 of it)."""
    g = load_golden(name)
    lin = O.mel2linear(g["mel"], g["feat_len"], g["mel_weight"], **M2L_CASES[name])
    scale = np.maximum(g["lin"].max(axis=-1, keepdims=True), 1e-10)
    assert_close(lin / scale, g["lin"] / scale, rtol=0, atol=1e-5, what=name)
    # silence rule: only utterances shorter than the longest are zeroed (-> clamp floor)
    fl = g["feat_len"]
    for i, n in enumerate(fl):
        if n != fl.max():
            assert (lin[i, n:] == 1e-10).all()


GL_CASES = {
    "gl_lin_512": dict(n_fft=512, hop=160, win=400, mag_spec=False, preemphasis=None),
    "gl_lin_pre_400": dict(n_fft=400, hop=160, win=400, mag_spec=True, preemphasis=0.97),
    "gl_mel_tts_1024": dict(n_fft=1024, hop=256, win=1024, mag_spec=False, preemphasis=0.97),
}


def hann_periodic(n):
    return (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)).astype(np.float32)


@pytest.mark.parametrize("name", sorted(GL_CASES))
def test_griffinlim_recover(name):
    """Griffin-Lim (32 iterations with momentum) started from the reference's own random phase.  The iteration
    amplifies rounding differences at the few bins whose phase is ill-conditioned, so two correct
    implementations (fp64 here, fp32 in the reference) agree to ~1e-4 of the peak in rms with isolated samples
    up to ~1e-3: asserted as rms <= 5e-4 peak and max <= 5e-3 peak (sfe_testutil.assert_gl_close)."""
    g = load_golden(name)
    c = GL_CASES[name]
    spec = g["feat"]
    if name.startswith("gl_mel"):
        W = O.mel_filterbank(22050, 1024, 80, 0.0, 8000.0).astype(np.float32)
        spec = O.mel2linear(spec, g["feat_len"], W, logging=True, log_base=None)
    wav, wav_len = O.speech2linear_recover(spec, g["feat_len"], c["n_fft"], c["hop"], c["win"], hann_periodic(c["win"]),
                                           c["mag_spec"], c["preemphasis"], g["angles"])
    ref = g["wav"].reshape(wav.shape)
    assert np.array_equal(wav_len, g["wav_len"])
    assert_gl_close(wav, ref, what=name)
    for i, n in enumerate(wav_len):
        assert (wav[i, n:] == 0).all()


def test_preemphasis_known_answer():
    # [1,2,3,4,5] -> [1,1.03,1.06,1.09,1.12]; y[len:] = 0
    x = np.array([[1, 2, 3, 4, 5.0]])
    y = O.preemphasis(x, np.array([5]), 0.97)
    assert np.allclose(y, [[1, 1.03, 1.06, 1.09, 1.12]])
    y = O.preemphasis(x, np.array([3]), 0.97)
    assert np.allclose(y, [[1, 1.03, 1.06, 0, 0]])


def test_seconds_vs_samples():
    # App. A #2
    assert O.to_samples(0.025, 16000) == 400 and O.to_samples(0.010, 16000) == 160
    assert O.to_samples(0.05, 22050) == 1102 and O.to_samples(0.0125, 22050) == 275
    assert O.to_samples(256, 22050) == 256
    assert O.FrontendCfg(hop_length=0.01, win_length=0.025).n_fft == 400  # App. A #3


def test_utterance_norm_chain():
    g = load_golden("chain_c2")
    y = O.feat_norm_utterance(g["feat"].astype(np.float64), g["feat_len"])
    assert_close(y, g["normed"], what="utterance norm")
    # SpecAugment on the CPU generator: masks bit-exact
    rng = O.TorchCpuRandint(int(g["seed"]))
    cfg = O.SpecAugCfg(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)
    B, F, D = g["normed"].shape
    d = O.specaug_draw(rng, B, D, g["feat_len"], cfg)
    out = O.specaug_apply(g["normed"], g["feat_len"], d, feat_norm=True)
    assert np.array_equal(out, g["augmented"])


SA_CASES = {
    "sa_recipe": dict(time_warp=False, freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5),
    "sa_intwidth": dict(time_warp=False, freq_mask_width=[3, 9], freq_mask_num=3, time_mask_width=20, time_mask_num=4),
    "sa_nonorm": dict(time_warp=False, feat_norm=False),
    "sa_freq_only": dict(time_warp=False, time_mask=False),
    "sa_time_only": dict(time_warp=False, freq_mask=False, time_mask_width=[0.02, 0.2]),
    "sa_warp": dict(time_warp=True),
    "sa_warp_only": dict(time_warp=True, freq_mask=False, time_mask=False, time_warp_window=8),
    "sa_warp_bilinear": dict(time_warp=True, time_warp_mode="bilinear"),
    "sa_short": dict(),
}


@pytest.mark.parametrize("name", sorted(SA_CASES))
def test_specaug(name):
    g = load_golden(name)
    cfg = O.SpecAugCfg(**SA_CASES[name])
    x, xl = g["x"], g["x_len"]
    d = O.specaug_draw(O.TorchCpuRandint(int(g["seed"])), x.shape[0], x.shape[2], xl, cfg)
    out = O.specaug_apply(x, xl, d, feat_norm=cfg.feat_norm)
    ref = g["y"]
    # mask pattern bit-exact: masked positions are where the reference holds the fill value
    mask = O.specaug_mask(*x.shape, d)
    fill = 0.0 if cfg.feat_norm else None
    if fill is not None and (cfg.freq_mask or cfg.time_mask):
        assert np.array_equal(ref == 0.0, mask)
    if d.warp is None:
        if cfg.feat_norm:
            assert np.array_equal(out, ref)
        else:
            assert np.array_equal(mask, ref != x) or np.allclose(out, ref, atol=1e-6)
            assert np.allclose(out, ref, atol=1e-6)
    else:
        assert_close(out, ref, rtol=1e-5, atol=1e-5, what=name)
        assert np.array_equal(out == 0.0, ref == 0.0) or not (cfg.freq_mask or cfg.time_mask)


def test_specaug_short_skips_warp():
    g = load_golden("sa_short")
    d = O.specaug_draw(O.TorchCpuRandint(int(g["seed"])), 3, 80, g["x_len"], O.SpecAugCfg())
    assert d.warp is None  # minlen 11 <= 2*5+1


def test_global_norm_sequence():
    g = load_golden("gnorm_seq")
    st = O.RunningStats(max_epoch_num=4)
    for i, ep in enumerate(g["epochs"]):
        ep = None if ep < 0 else int(ep)
        y = O.feat_norm_global([(g[f"x{i}"].astype(np.float64), g[f"len{i}"])], st, ep)[0]
        assert st.batch == int(g[f"batch{i}"][0]), f"step {i}"
        assert_close(st.mean, g[f"mean{i}"], rtol=1e-5, atol=1e-6, what=f"mean{i}")
        assert_close(st.std, g[f"std{i}"], rtol=1e-5, atol=1e-6, what=f"std{i}")
        assert_close(y, g[f"y{i}"], what=f"y{i}")
    y = O.feat_norm_global([(g["x0"].astype(np.float64), g["len0"])], st, None, training=False)[0]
    assert_close(y, g["y_eval"], what="eval")
    assert list(g["state_keys"]) == ["global_batch", "global_mean", "global_std"]


def test_global_norm_two_rank_emulation_equals_single():
    g = load_golden("gnorm_seq")
    x, l = g["x0"].astype(np.float64), g["len0"]
    one = O.feat_norm_global([(x, l)], O.RunningStats(), None)[0]
    # rank r of W takes batch[W-1-r::W] (iterator/abs.py:231-236)
    shards = [(x[1::2], l[1::2]), (x[0::2], l[0::2])]
    two = O.feat_norm_global(shards, O.RunningStats(), None)
    assert np.allclose(two[0], one[1::2], rtol=1e-12) and np.allclose(two[1], one[0::2], rtol=1e-12)


def test_batch_norm():
    g = load_golden("bnorm")
    x, l = g["x"].astype(np.float64), g["x_len"]
    assert_close(O.feat_norm_batch([(x, l)], training=True)[0], g["y_train"], what="train")
    assert_close(O.feat_norm_batch([(x, l)], training=False)[0], g["y_eval"], what="eval")


def test_norm_flags_and_1d():
    g = load_golden("unorm_meanonly")
    assert_close(O.feat_norm_utterance(g["x"].astype(np.float64), g["x_len"], std_norm=False), g["y"])
    g = load_golden("unorm_stdonly")
    assert_close(O.feat_norm_utterance(g["x"].astype(np.float64), g["x_len"], mean_norm=False), g["y"])
    g = load_golden("norm_1d")
    x = g["x"].astype(np.float64)[..., None]
    assert_close(O.feat_norm_utterance(x, g["x_len"])[..., 0], g["y_utt"])
    assert_close(O.feat_norm_global([(x, g["x_len"])], O.RunningStats(), 1)[0][..., 0], g["y_global"])


def test_single_frame_utterance_std_is_nan():
    # App. A #13: unbiased std of one frame is NaN in the reference
    x = np.random.default_rng(0).normal(size=(2, 5, 4))
    m, s = O.utterance_stats(x, [5, 1])
    assert np.isnan(s[1]).all() and np.isfinite(s[0]).all()


def test_torch_randint_restatement_matches_torch():
    torch = pytest.importorskip("torch")
    for seed in (0, 7, 1234, 2**31 + 5):
        torch.manual_seed(seed)
        rng = O.TorchCpuRandint(seed)
        for lo, hi, size in [(0, 31, (64, 2)), (0, 50, (64, 2)), (6, 100, (1,)), (3, 2000, (5, 7))]:
            assert np.array_equal(torch.randint(lo, hi, size).numpy(), rng.randint(lo, hi, size))


def test_staged_reference_modules_equal_the_port():
    """bench.py's CPU arm: oracle/_ref (the reference's own modules, staged by oracle/build_ref.py) and the port
    (oracle/ref_port_torch.py) must be interchangeable bit for bit, masks included."""
    import torch

    from oracle import build_ref, ref_modules
    from oracle import ref_port_torch as port
    if not build_ref.build(quiet=True):
        pytest.skip("no /root/reference and nothing staged under oracle/_ref")
    import bench
    for name, lens in (("c2", [50000, 40000, 32123, 8000]), ("c3", [44100, 22050]), ("c4", [48000, 30000])):
        w = dict(bench.WORKLOADS[name])
        fc = w["frontend"]
        sr = fc.get("sr", 16000)
        win = int(fc["win_length"] * sr) if isinstance(fc["win_length"], float) else fc["win_length"]
        args = (sr, fc.get("n_fft") or win, fc["n_mels"], fc.get("fmin", 0.0), fc.get("fmax"))
        wav = bench.make_batch(w, lens, 3, "cpu")
        wl = torch.tensor(lens)
        bench.cpu_pass(dict(w), wav.clone(), wl.clone(), ref_modules.mel_weight(*args), ref_modules)  # (builds the modules: nn.Conv1d init draws)
        torch.manual_seed(1)
        a = bench.cpu_pass(dict(w), wav.clone(), wl.clone(), ref_modules.mel_weight(*args), ref_modules)
        torch.manual_seed(1)
        b = bench.cpu_pass(dict(w), wav.clone(), wl.clone(), port.mel_weight(*args), port)
        assert torch.equal(a.nan_to_num(), b.nan_to_num()), name

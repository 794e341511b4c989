"""GPU parity: the CUDA path (through the module API -> C ABI) against the golden fixtures from
the reference and against the CPU oracle on seeded inputs.  Tolerance (north star): within
1e-4 relative / 1e-3 absolute in fp32; SpecAugment masks bit-exact."""
import numpy as np
import pytest
import torch

from oracle import frontend_np as O
from sfe_testutil import assert_close, assert_gl_close, load_golden
from test_oracle_golden import FE_CASES, SA_CASES

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from speechain_b200.host import ops
    from speechain_b200.module.acoustic_frontend import AcousticFrontend
    from speechain_b200.module.augment.specaug import SpecAugment
    from speechain_b200.module.frontend.delta_feat import DeltaFeature
    from speechain_b200.module.frontend.linear2mel import LinearSpec2MelSpec
    from speechain_b200.module.frontend.speech2linear import Speech2LinearSpec
    from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
    from speechain_b200.module.norm.feat_norm import FeatureNormalization

DEV = "cuda:0"


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def test_native_library_is_loaded():
    from speechain_b200 import _lib
    lib = _lib.load()
    assert lib.sfe_abi_version() == 1
    assert 512 in _lib.supported_nfft() and 400 in _lib.supported_nfft() and 1024 in _lib.supported_nfft()


# ------------------------------------------------------------------------------------------
# frontend vs golden (reference outputs)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(FE_CASES))
def test_speech2mel_golden(name):
    g = load_golden(name)
    m = Speech2MelSpec(**FE_CASES[name]).to(DEV)
    wl = cu(g["wav_len"])
    wl_before = wl.clone()
    out = m(cu(g["wav"])[..., None], wl)
    feat, flen = out[0], out[1]
    assert torch.equal(wl, wl_before), "speech_len must not be mutated"
    assert np.array_equal(flen.cpu().numpy(), g["feat_len"])
    assert_close(feat.cpu().numpy(), g["feat"], what=name)
    if "energy" in g:
        assert_close(out[2].cpu().numpy(), g["energy"], what=name + " energy")
        assert np.array_equal(out[3].cpu().numpy(), g["feat_len"])
    # bit-identical filterbank buffer (checkpoint compatibility)
    assert np.array_equal(m.linear2mel.mel_fbanks.weight.cpu().numpy(), g["mel_weight"])


@pytest.mark.parametrize("name", ["fe_c1", "fe_c2_400", "fe_c3"])
def test_speech2mel_2d_input_and_unaligned_view(name):
    g = load_golden(name)
    m = Speech2MelSpec(**FE_CASES[name]).to(DEV)
    wav = g["wav"]
    # (B, T) input, and a view whose base pointer is only 4-byte aligned with an odd pitch
    big = torch.zeros(wav.shape[0], wav.shape[1] + 3, device=DEV)
    big[:, 1:1 + wav.shape[1]] = cu(wav)
    view = big[:, 1:1 + wav.shape[1]]
    out = m(view, cu(g["wav_len"]))
    assert_close(out[0].cpu().numpy(), g["feat"], what=name + " view")


@pytest.mark.parametrize("name,conf", [
    ("lin_mag_log", dict(sr=22050, win_length=1024, hop_length=256, mag_spec=True, logging=True, clamp=1e-5)),
    ("lin_pow", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, return_energy=True)),
    ("lin_pow_log10", dict(sr=16000, win_length=400, hop_length=160, preemphasis=0.97, logging=True, log_base=10.0)),
    ("lin_twosided", dict(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False, return_energy=True)),
    ("lin_twosided_mag_log", dict(sr=16000, win_length=400, hop_length=160, onesided=False, mag_spec=True, logging=True,
                                  clamp=1e-5, return_energy=True)),
])
def test_speech2linear_golden(name, conf):
    g = load_golden(name)
    m = Speech2LinearSpec(**conf).to(DEV)
    out = m(cu(g["wav"])[..., None], cu(g["wav_len"]))
    spec = out[0].cpu().numpy()
    assert np.array_equal(out[1].cpu().numpy(), g["feat_len"])
    if name == "lin_pow_log10":
        # single bins at the fp32 noise floor: compare in the power domain (see test_oracle_golden.py)
        p, pref = 10.0 ** spec.astype(np.float64), 10.0 ** g["spec"].astype(np.float64)
        assert_close(p, pref, rtol=1e-4, atol=1e-6 * pref.max(), what=name)
    elif name in ("lin_mag_log", "lin_twosided_mag_log"):
        p, pref = np.exp(spec.astype(np.float64)), np.exp(g["spec"].astype(np.float64))
        assert_close(p, pref, rtol=1e-4, atol=1e-5 * pref.max(), what=name)
    else:
        assert_close(spec, g["spec"], what=name)
    if "energy" in g:
        assert_close(out[2].cpu().numpy(), g["energy"], what=name + " energy")


def test_twosided_spectrogram_with_waveform_normalisation_and_no_recovery():
    """onesided=False on the rare pre_stft_norm path against the oracle, and recover() refusing all-n_fft-bin input."""
    conf = dict(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False, return_energy=True, logging=True,
                pre_stft_norm="min_max")
    g = torch.Generator().manual_seed(77)
    lens = np.array([4000, 2600, 900])
    wav = torch.randn(3, 4000, generator=g) * 0.1
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    m = Speech2LinearSpec(**conf).to(DEV)
    spec, flen, energy, _ = m(wav.to(DEV)[..., None], cu(lens))
    cfg = O.FrontendCfg(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False, return_energy=True,
                        lin_logging=True, pre_stft_norm="min_max")
    ref, rlen, renergy = O.speech2linear(wav.numpy().astype(np.float64), lens, cfg)
    assert spec.shape[-1] == 512 and np.array_equal(flen.cpu().numpy(), rlen)
    p, pref = np.exp(spec.cpu().numpy().astype(np.float64)), np.exp(ref)
    assert_close(p, pref, rtol=1e-4, atol=1e-6 * pref.max(), what="two-sided prenorm")
    assert_close(energy.cpu().numpy(), renergy, what="two-sided prenorm energy")
    with pytest.raises(ValueError):
        Speech2LinearSpec(sr=16000, win_length=400, hop_length=160, n_fft=512, onesided=False).to(DEV).recover(spec, flen)


def test_linear2mel_golden():
    g = load_golden("l2m")
    m = LinearSpec2MelSpec(n_fft=512, n_mels=80).to(DEV)
    mel, _ = m(cu(g["spec"]), cu(np.array([30, 20])))
    assert_close(mel.cpu().numpy(), g["mel"], what="l2m")


@pytest.mark.parametrize("name,conf", [
    ("m2l_log10", dict(n_fft=512, n_mels=80)),
    ("m2l_ln_tts", dict(n_fft=1024, n_mels=80, sr=22050, fmin=0, fmax=8000, clamp=1e-5, log_base=None)),
    ("m2l_nolog", dict(n_fft=512, n_mels=80, logging=False)),
])
def test_mel2linear_recover_golden(name, conf):
    """LinearSpec2MelSpec.recover against the reference's lstsq output (relative to each frame's largest bin:
    the least-squares solution cancels inside a frame) and against the fp64 oracle."""
    g = load_golden(name)
    m = LinearSpec2MelSpec(**conf).to(DEV)
    assert np.array_equal(m.mel_fbanks.weight.cpu().numpy(), g["mel_weight"])
    for flen in (cu(g["feat_len"]), torch.from_numpy(g["feat_len"])):   # device or host lengths
        lin = m.recover(cu(g["mel"]), flen).cpu().numpy()
        scale = np.maximum(g["lin"].max(axis=-1, keepdims=True), 1e-10)
        assert_close(lin / scale, g["lin"] / scale, rtol=0, atol=1e-5, what=name)
        ref = O.mel2linear(g["mel"], g["feat_len"], g["mel_weight"], logging=conf.get("logging", True),
                           log_base=conf.get("log_base", 10.0))
        assert_close(lin / scale, ref / scale, rtol=0, atol=1e-5, what=name + " vs oracle")
        assert lin.min() >= 1e-10
        fl = g["feat_len"]
        for i, n in enumerate(fl):
            if n != fl.max():
                assert (lin[i, n:] == np.float32(1e-10)).all()


def test_mel2linear_roundtrip_full_size():
    """Size-independent property at a C2-like size: mel(recover(mel)) == mel, because W pinv(W) = I."""
    torch.manual_seed(5)
    m = LinearSpec2MelSpec(n_fft=512, n_mels=80).to(DEV)
    mel = torch.randn(16, 1200, 80, device=DEV) * 2 - 3      # log10 mel energies
    flen = torch.full((16,), 1200, dtype=torch.int64, device=DEV)
    lin = m.recover(mel, flen)
    assert lin.shape == (16, 1200, 257)
    # the clamp at 1e-10 only touches bins the unconstrained solution drives negative: undo it for the check.
    # A frame spans many decades (10^(N(-3, 2))), so every tolerance is relative to the frame's own terms.
    pinv_t = m._pinv_t(mel.device).double()
    e = torch.pow(10.0, mel.double())
    raw = e @ pinv_t
    back = raw @ m.mel_fbanks.weight.double().T
    frame_max = e.amax(dim=-1, keepdim=True).cpu().numpy()
    assert_close(back.cpu().numpy(), e.cpu().numpy(), rtol=1e-6, atol=1e-6 * frame_max, what="W pinv(W) = I")
    # fp32 accumulation of D terms: |err| <= D eps sum_d |e_d pinv[d, k]|
    term_sum = (e @ pinv_t.abs()).cpu().numpy()
    assert_close(lin.cpu().numpy(), raw.clamp(min=1e-10).cpu().numpy(), rtol=1e-4, atol=3e-6 * term_sum + 1e-10,
                 what="kernel vs fp64 product")


GL_GPU_CASES = {
    "gl_lin_512": ("Speech2LinearSpec", dict(hop_length=160, win_length=400, n_fft=512, sr=16000)),
    "gl_lin_pre_400": ("Speech2LinearSpec", dict(hop_length=160, win_length=400, sr=16000, preemphasis=0.97, mag_spec=True)),
    "gl_mel_tts_1024": ("Speech2MelSpec", dict(n_mels=80, hop_length=256, win_length=1024, sr=22050, fmin=0, fmax=8000,
                                             clamp=1e-5, log_base=None, preemphasis=0.97)),
}


@pytest.mark.parametrize("name", sorted(GL_GPU_CASES))
def test_griffinlim_recover_golden(name):
    """recover() (Griffin-Lim + inverse pre-emphasis) against the reference's CPU output: the module draws the
    initial phase from the global CPU generator exactly like torchaudio, so seeding reproduces the reference."""
    g = load_golden(name)
    cls, conf = GL_GPU_CASES[name]
    m = globals()[cls](**conf).to(DEV)
    torch.manual_seed(int(g["seed"]))
    wav, wav_len = m.recover(cu(g["feat"]), cu(g["feat_len"]))
    assert tuple(wav.shape) == tuple(g["wav"].shape)          # (B, L, 1) with pre-emphasis, (B, L) without
    assert np.array_equal(wav_len.cpu().numpy(), g["wav_len"])
    assert_gl_close(wav.cpu().numpy(), g["wav"], what=name)
    w = wav.cpu().numpy().reshape(len(g["wav_len"]), -1)
    for i, n in enumerate(g["wav_len"]):
        assert (w[i, n:] == 0).all()


def test_griffinlim_single_iteration_matches_oracle():
    """One STFT/ISTFT round trip at every compiled FFT size against the fp64 oracle (tight tolerance: no chaos yet)."""
    from speechain_b200.host.ops import FrontendPlan
    rng = np.random.default_rng(7)
    for n_fft, win, hop in [(128, 128, 32), (200, 200, 50), (256, 200, 64), (320, 320, 80), (400, 400, 160), (480, 400, 120),
                            (512, 400, 160), (640, 640, 160), (800, 800, 200), (1024, 1024, 256), (2048, 1200, 300)]:
        B, F, K = 2, 12, n_fft // 2 + 1
        window = torch.hann_window(win)
        plan = FrontendPlan(n_fft=n_fft, win_length=win, hop_length=hop, window=window, device=DEV)
        spec = rng.random((B, F, K)).astype(np.float32) ** 2 + 1e-3
        ang = (rng.random((B, F, K)) + 1j * rng.random((B, F, K))).astype(np.complex64)
        flen = np.array([F, F - 3])
        for n_iter in (0, 2):
            wav, wl = plan.griffinlim(cu(spec), cu(flen), n_iter=n_iter, momentum=0.99, init_angles=torch.from_numpy(ang))
            ref = O.griffinlim(spec, n_fft, hop, win, window.numpy(), 2.0, ang, n_iter=n_iter)
            for i, n in enumerate((flen - 1) * hop):
                ref[i, n:] = 0
            assert np.array_equal(wl.cpu().numpy(), (flen - 1) * hop)
            assert_close(wav.cpu().numpy(), ref, rtol=0, atol=2e-5 * np.abs(ref).max(), what=f"n_fft {n_fft} n_iter {n_iter}")


def test_griffinlim_spectral_convergence_full_size():
    """Size-independent property at a TTS-batch size: the recovered waveform's spectrogram approaches the input
    magnitudes (spectral convergence well below the random-phase start) and the output is finite."""
    conf = dict(hop_length=256, win_length=1024, sr=22050, mag_spec=True)
    fe = Speech2LinearSpec(**conf).to(DEV)
    torch.manual_seed(3)
    t = torch.arange(22050 * 4, device=DEV) / 22050
    wav = torch.stack([0.4 * torch.sin(2 * np.pi * (110 * (i + 1)) * t) * torch.sin(2 * np.pi * 3 * t) for i in range(16)])[..., None]
    wl = torch.full((16,), wav.size(1), dtype=torch.int64, device=DEV)
    spec, fl = fe(wav, wl)
    rec, rl = Speech2LinearSpec(**conf).to(DEV).recover(spec, fl)
    assert torch.isfinite(rec).all()
    spec2, _ = fe(rec[..., None] if rec.dim() == 2 else rec, rl)
    F = min(spec.size(1), spec2.size(1))
    sc = (torch.linalg.norm(spec[:, :F] - spec2[:, :F]) / torch.linalg.norm(spec[:, :F])).item()
    assert sc < 0.1, sc


def test_delta_matches_oracle():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(3, 37, 80)).astype(np.float32)
    for order in (1, 2):
        y, _ = DeltaFeature(delta_order=order, delta_N=2).to(DEV)(cu(x), cu(np.array([37, 30, 12])))
        assert_close(y.cpu().numpy(), O.delta_feature(x.astype(np.float64), order, 2), rtol=1e-5, atol=1e-5)


def test_unsupported_inputs_raise():
    with pytest.raises(ValueError):
        Speech2MelSpec(n_mels=80, hop_length=100, win_length=300, n_fft=300).to(DEV)(
            torch.zeros(1, 4000, 1, device=DEV), torch.tensor([4000], device=DEV))
    # a log floor below the smallest normal float would not be exact on the SFU path: refused, not approximated
    with pytest.raises(ValueError):
        Speech2MelSpec(n_mels=80, hop_length=160, win_length=400, n_fft=512, clamp=1e-40).to(DEV)(
            torch.zeros(1, 4000, 1, device=DEV), torch.tensor([4000], device=DEV))
    m = Speech2MelSpec(n_mels=80, hop_length=160, win_length=400, n_fft=512)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4000, 2, device=DEV), torch.tensor([4000, 4000], device=DEV))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 4000, 1), torch.tensor([4000, 4000]))  # CPU tensors: no CPU path


# ------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------
def test_utterance_norm_and_specaug_chain_golden():
    g = load_golden("chain_c2")
    feat, fl = cu(g["feat"]), cu(g["feat_len"])
    un = FeatureNormalization(input_size=80, norm_type="utterance").to(DEV)
    y, _ = un(feat, fl)
    assert_close(y.cpu().numpy(), g["normed"], what="utterance norm")
    assert torch.equal(feat, cu(g["feat"])), "input must not be modified"
    sa = SpecAugment(input_size=80, time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                     time_mask_num=2, feat_norm=True).to(DEV)
    sa.train()
    torch.manual_seed(int(g["seed"]))
    a, _ = sa(cu(g["normed"]), fl)
    assert np.array_equal(a.cpu().numpy(), g["augmented"]), "masking an identical input must be bit-exact"


def test_fused_chain_equals_reference_chain():
    gfe, gch = load_golden("chain_c2_fe"), load_golden("chain_c2")
    chain = AcousticFrontend(frontend=dict(conf=FE_CASES["chain_c2_fe"]), normalize=dict(norm_type="utterance"),
                             specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                                          time_mask_num=2)).to(DEV)
    chain.train()
    torch.manual_seed(int(gch["seed"]))
    y, yl = chain(cu(gfe["wav"])[..., None], cu(gfe["wav_len"]))
    got, ref = y.cpu().numpy(), gch["augmented"]
    assert np.array_equal(yl.cpu().numpy(), gch["feat_len"])
    assert np.array_equal(got == 0.0, ref == 0.0), "mask pattern must be bit-exact"
    assert_close(got, ref, what="fused chain")
    chain.eval()
    y, _ = chain(cu(gfe["wav"])[..., None], cu(gfe["wav_len"]))
    assert_close(y.cpu().numpy(), gch["normed"], what="fused chain eval")


def test_global_norm_sequence_golden():
    g = load_golden("gnorm_seq")
    gn = FeatureNormalization(input_size=80, norm_type="global").to(DEV)
    gn.train()
    for i, ep in enumerate(g["epochs"]):
        ep = None if ep < 0 else int(ep)
        y, _ = gn(cu(g[f"x{i}"]), cu(g[f"len{i}"]), epoch=ep)
        assert int(gn.global_batch.item()) == int(g[f"batch{i}"][0]), f"step {i}"
        assert_close(gn.global_mean.cpu().numpy(), g[f"mean{i}"], rtol=1e-5, atol=1e-5, what=f"mean{i}")
        assert_close(gn.global_std.cpu().numpy(), g[f"std{i}"], rtol=1e-5, atol=1e-5, what=f"std{i}")
        assert_close(y.cpu().numpy(), g[f"y{i}"], what=f"y{i}")
    assert sorted(gn.state_dict().keys()) == list(g["state_keys"])
    gn.eval()
    y, _ = gn(cu(g["x0"]), cu(g["len0"]))
    assert_close(y.cpu().numpy(), g["y_eval"], what="eval")
    # checkpoint round trip through the lazy loader (feat_norm.py:609-632)
    fresh = FeatureNormalization(input_size=80, norm_type="global").to(DEV)
    fresh.load_state_dict(gn.state_dict(), strict=True)
    fresh.eval()
    y2, _ = fresh(cu(g["x0"]), cu(g["len0"]))
    assert torch.equal(y, y2)


def test_batch_norm_golden():
    g = load_golden("bnorm")
    bn = FeatureNormalization(input_size=80, norm_type="batch").to(DEV)
    bn.train()
    y, _ = bn(cu(g["x"]), cu(g["x_len"]))
    assert_close(y.cpu().numpy(), g["y_train"], what="train")
    bn.eval()
    y, _ = bn(cu(g["x"]), cu(g["x_len"]))
    assert_close(y.cpu().numpy(), g["y_eval"], what="eval")


def test_norm_flags_and_1d_golden():
    g = load_golden("unorm_meanonly")
    y, _ = FeatureNormalization(input_size=80, norm_type="utterance", std_norm=False).to(DEV)(cu(g["x"]), cu(g["x_len"]))
    assert_close(y.cpu().numpy(), g["y"])
    g = load_golden("unorm_stdonly")
    y, _ = FeatureNormalization(input_size=80, norm_type="utterance", mean_norm=False).to(DEV)(cu(g["x"]), cu(g["x_len"]))
    assert_close(y.cpu().numpy(), g["y"])
    g = load_golden("norm_1d")
    y, _ = FeatureNormalization(input_size=1, norm_type="utterance").to(DEV)(cu(g["x"]), cu(g["x_len"]))
    assert y.dim() == 2
    assert_close(y.cpu().numpy(), g["y_utt"])
    gn = FeatureNormalization(input_size=1, norm_type="global").to(DEV)
    gn.train()
    y, _ = gn(cu(g["x"]), cu(g["x_len"]), epoch=1)
    assert_close(y.cpu().numpy(), g["y_global"])


def test_group_norm_golden():
    g = load_golden("grpnorm_tensor")
    m = FeatureNormalization(input_size=80, norm_type="group").to(DEV)
    m.train()
    for i in range(2):
        y, _ = m(cu(g[f"x{i}"]), cu(g[f"len{i}"]), group_ids=cu(g[f"ids{i}"]), epoch=1)
        assert_close(y.cpu().numpy(), g[f"y{i}"], what=f"group y{i}")
    for k in [k for k in g if k.startswith("buf_")]:
        assert_close(m.state_dict()[k[4:]].cpu().numpy(), g[k], rtol=1e-5, atol=1e-5, what=k)
    m.eval()
    y, _ = m(cu(g["x_eval"]), cu(g["len_eval"]), group_ids=cu(g["ids_eval"]))
    assert_close(y.cpu().numpy(), g["y_eval"], what="group eval")
    g = load_golden("grpnorm_str")
    m = FeatureNormalization(input_size=80, norm_type="group").to(DEV)
    m.train()
    y0, _ = m(cu(g["x0"]), cu(g["len0"]), group_ids="real", epoch=1)
    y1, _ = m(cu(g["x1"]), cu(g["len1"]), group_ids="syn", epoch=1)
    assert_close(y0.cpu().numpy(), g["y0"])
    assert_close(y1.cpu().numpy(), g["y1"])
    assert sorted(m.state_dict().keys()) == list(g["state_keys"])


def test_single_frame_utterance_gives_nan_std():
    x = torch.randn(2, 5, 8, device=DEV)
    y, _ = FeatureNormalization(input_size=8, norm_type="utterance").to(DEV)(x, torch.tensor([5, 1], device=DEV))
    assert torch.isnan(y[1]).all() and torch.isfinite(y[0]).all()


# ------------------------------------------------------------------------------------------
# SpecAugment
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(SA_CASES))
def test_specaug_golden(name):
    g = load_golden(name)
    sa = SpecAugment(input_size=80, **SA_CASES[name]).to(DEV)
    sa.train()
    torch.manual_seed(int(g["seed"]))
    x = cu(g["x"])
    y, _ = sa(x, cu(g["x_len"]))
    got, ref = y.cpu().numpy(), g["y"]
    assert torch.equal(x, cu(g["x"])), "input must not be modified"
    warped = sa.time_warp and int(g["x_len"].min()) > 2 * sa.time_warp_window + 1
    if sa.feat_norm and (sa.freq_mask or sa.time_mask):
        assert np.array_equal(got == 0.0, ref == 0.0), "mask pattern must be bit-exact"
    if not warped and sa.feat_norm:
        assert np.array_equal(got, ref)
    else:
        assert_close(got, ref, rtol=1e-5, atol=2e-5, what=name)


def test_specaug_latent_assert_is_kept():
    sa = SpecAugment(time_warp=False).to(DEV)  # input_size=None (specaug.py:135-141)
    sa.train()
    with pytest.raises(AssertionError):
        sa(torch.zeros(2, 50, 80, device=DEV), torch.tensor([50, 40], device=DEV))


# ------------------------------------------------------------------------------------------
# seeded inputs vs the oracle at BASELINE sizes (C1 / C2 / C3), plus size-independent properties
# ------------------------------------------------------------------------------------------
def _wave(cfg_id, lens, zero_seg=True):
    g = torch.Generator().manual_seed(1234 + cfg_id)
    wav = 0.1 * torch.randn(len(lens), max(lens), generator=g)
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    if zero_seg:
        wav[0, 1000:9000] = 0  # exact-zero segment: exercises the clamp floor
    return wav


def test_c1_full_size_vs_oracle():
    lens = [240000] * 32
    wav = _wave(1, lens)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000)
    feat, flen = Speech2MelSpec(**conf).to(DEV)(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    assert feat.shape == (32, 1501, 80)
    sub = [0, 13, 31]
    ref, rl, _ = O.speech2mel(wav[sub].numpy().astype(np.float64), [lens[i] for i in sub], O.FrontendCfg(**conf))
    assert_close(feat[sub].cpu().numpy(), ref, what="C1")


def test_c2_full_size_vs_oracle_and_properties():
    g = torch.Generator().manual_seed(1236)
    lens = sorted(torch.randint(32000, 320001, (64,), generator=g).tolist(), reverse=True)
    wav = _wave(2, lens)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"),
                             specaug=dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05,
                                          time_mask_num=2)).to(DEV)
    chain.train()
    torch.manual_seed(1234)
    y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    cfg = O.FrontendCfg(**conf)
    flen = O.feat_lengths(np.array(lens), 512, 160, True)
    assert np.array_equal(yl.cpu().numpy(), flen)
    d = O.specaug_draw(O.TorchCpuRandint(1234), 64, 80, flen, O.SpecAugCfg(time_warp=False))
    mask = O.specaug_mask(64, y.shape[1], 80, d)
    got = y.cpu().numpy()
    assert (got[mask] == 0.0).all(), "every masked element must hold the fill value"
    # un-masked elements are non-zero, except that among ~1e7 values one may equal its utterance mean
    # exactly in fp32 ((x - mean) / std == 0.0): allow a handful of such natural zeros
    assert np.count_nonzero((got == 0.0) & ~mask) <= 4, "mask pattern must match the reference draw sequence"
    # the batch-level reflect padding means utterance 0 (len == T_max) differs from a solo run only
    # in its last frame; check three utterances against the oracle run on the *padded batch rows*
    sub = [0, 20, 63]
    T = max(lens)
    ref, _, _ = O.speech2mel(wav[sub].numpy().astype(np.float64), [lens[i] for i in sub], cfg)
    ref = O.feat_norm_utterance(ref, flen[sub])
    ref = np.where(mask[sub], 0.0, ref)
    assert_close(got[sub], ref, what="C2")
    # property: unmasked valid frames of every utterance have zero mean / unit std before masking
    chain.eval()
    z, _ = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    z = z.cpu().numpy().astype(np.float64)
    for i in (0, 31, 63):
        v = z[i, :flen[i]]
        assert np.abs(v.mean(0)).max() < 2e-4 and np.abs(v.std(0, ddof=1) - 1).max() < 2e-4


def test_c3_full_size_vs_oracle():
    lens = [220500] * 128
    wav = _wave(3, lens)
    conf = dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000, log_base=None,
                clamp=1e-5, return_energy=True)
    feat, flen, en, _ = Speech2MelSpec(**conf).to(DEV)(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    assert feat.shape == (128, 862, 80) and en.shape == (128, 862)
    sub = [0, 77]
    ref, _, ren = O.speech2mel(wav[sub].numpy().astype(np.float64), [lens[i] for i in sub], O.FrontendCfg(**conf))
    assert_close(feat[sub].cpu().numpy(), ref, what="C3")
    assert_close(en[sub].cpu().numpy(), ren, what="C3 energy")


def test_linearity_and_parseval_properties():
    # |STFT|^2 of a*x is a^2 * |STFT|^2 of x; energy^2 == sum_k P[k] (one-sided) — size independent
    lens = [48000, 40000]
    wav = _wave(7, lens, zero_seg=False).to(DEV)
    m = Speech2LinearSpec(hop_length=160, win_length=400, n_fft=512, return_energy=True).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    p1, _, e1, _ = m(wav[..., None], wl)
    p2, _, e2, _ = m((2.0 * wav)[..., None], wl)
    assert torch.allclose(p2, 4.0 * p1, rtol=1e-5, atol=1e-9)
    assert torch.allclose(e1[:, :250] ** 2, p1[:, :250].sum(-1), rtol=1e-5)


def test_repeated_calls_are_deterministic_and_scheduler_rearms():
    lens = [64000, 50000, 33333]
    wav = _wave(9, lens).to(DEV)
    m = Speech2MelSpec(n_mels=80, hop_length=0.01, win_length=0.025, preemphasis=0.97).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    a, _ = m(wav[..., None], wl)
    for _ in range(3):
        b, _ = m(wav[..., None], wl)
        assert torch.equal(a, b)


@pytest.mark.parametrize("B", [1, 3, 300])
def test_work_distribution_edge_batches(B):
    """Unit scheduling of the fused kernel: fewer units than warps (B = 1, 3) and batches beyond the cost table
    (B = 300 > 256: uniform unit costs), ragged lengths incl. utterances shorter than a frame hop; both the
    plain frontend (padded rows written by the kernel) and the chain (padded rows left to the normalise pass)."""
    g = torch.Generator().manual_seed(100 + B)
    lens = sorted(torch.randint(100, 6000, (B,), generator=g).tolist(), reverse=True)
    lens[0] = 6000
    wav = _wave(40 + B, lens, zero_seg=False)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    cfg = O.FrontendCfg(**conf)
    wl = torch.tensor(lens, device=DEV)
    feat, flen = Speech2MelSpec(**conf).to(DEV)(wav.to(DEV)[..., None], wl)
    ref, rl, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, cfg)
    assert np.array_equal(flen.cpu().numpy(), rl)
    assert_close(feat.cpu().numpy(), ref, what=f"frontend B={B}")
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance")).to(DEV)
    chain.eval()
    y, _ = chain(wav.to(DEV)[..., None], wl)
    yn, rn = y.cpu().numpy(), O.feat_norm_utterance(ref, rl)
    assert np.array_equal(np.isnan(yn), np.isnan(rn)), "NaN pattern (single-frame utterances have a NaN std)"
    ok = rl >= 8  # a std estimated from a handful of frames amplifies fp32 rounding beyond the tolerance
    assert_close(yn[ok], rn[ok], what=f"chain B={B}")
    # energy + linear outputs take the path that enumerates padded units
    lin, ll, en, _ = Speech2LinearSpec(hop_length=160, win_length=400, n_fft=512, return_energy=True).to(DEV)(
        wav.to(DEV)[..., None], wl)
    cfg_l = O.FrontendCfg(n_mels=80, n_fft=512, win_length=400, hop_length=160, sr=16000, return_energy=True)
    rlin, _, ren = O.speech2linear(wav.numpy().astype(np.float64), lens, cfg_l)
    assert_close(en.cpu().numpy(), ren, what=f"energy B={B}")
    assert_close(lin.cpu().numpy(), rlin, rtol=1e-4, atol=1e-5 * float(np.abs(rlin).max()), what=f"linear B={B}")


def test_memory_beyond_each_length_is_never_read():
    # the kernel treats samples >= len as zero by construction: poison them with NaN (also odd lengths / pitch)
    for conf in (dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97),
                 dict(n_mels=80, win_length=0.025, hop_length=0.010),
                 dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, log_base=None, clamp=1e-5)):
        lens = [40001, 39999, 25600, 25607, 16161, 4000]
        clean = _wave(11, lens, zero_seg=False).to(DEV)
        dirty = clean.clone()
        for i, n in enumerate(lens):
            dirty[i, n:] = float("nan")
        m = Speech2MelSpec(**conf).to(DEV)
        wl = torch.tensor(lens, device=DEV)
        a, _ = m(clean[..., None], wl)
        b, _ = m(dirty[..., None], wl)
        assert torch.isfinite(b).all()
        assert torch.equal(a, b)


def test_ragged_upload_equals_padded_copy():
    lens = [30000, 22222, 9999, 1]
    host = _wave(12, lens, zero_seg=False)[..., None].contiguous().pin_memory()
    wl = torch.tensor(lens, dtype=torch.int64)
    dev = torch.full(host.shape, float("nan"), device=DEV)
    ops.upload_waveforms(host, wl, out=dev)
    for i, n in enumerate(lens):
        assert torch.equal(dev[i, :n].cpu(), host[i, :n])
        assert torch.isnan(dev[i, n:]).all()  # padding is not transferred
    m = Speech2MelSpec(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010).to(DEV)
    a, _ = m(dev, wl.to(DEV))
    b, _ = m(host.to(DEV), wl.to(DEV))
    assert torch.equal(a, b)


@pytest.mark.parametrize("norm_type", ["utterance", "global"])
def test_fused_chain_with_time_warp_vs_oracle(norm_type):
    # recipe-default SpecAugment (time_warp on): frontend -> normalise -> warp -> masks
    g = torch.Generator().manual_seed(4242)
    lens = sorted(torch.randint(16000, 64001, (6,), generator=g).tolist(), reverse=True)
    wav = _wave(21, lens, zero_seg=False)
    conf = dict(n_mels=80, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)  # n_fft = 400
    sconf = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type=norm_type), specaug=sconf).to(DEV)
    chain.train()
    torch.manual_seed(99)
    y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV), epoch=1)
    feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
    if norm_type == "utterance":
        ref = O.feat_norm_utterance(feat, flen)
    else:
        ref = O.feat_norm_global([(feat, flen)], O.RunningStats(), 1)[0]
    d = O.specaug_draw(O.TorchCpuRandint(99), len(lens), 80, flen, O.SpecAugCfg(**sconf))
    assert d.warp is not None
    ref = O.specaug_apply(ref, flen, d, feat_norm=True)
    got = y.cpu().numpy()
    mask = O.specaug_mask(*got.shape, d)
    assert (got[mask] == 0.0).all() and np.count_nonzero((got == 0.0) & ~mask) <= 2
    assert_close(got, ref, what=f"warp chain {norm_type}")


@pytest.mark.parametrize("mode", ["bicubic", "bilinear"])
@pytest.mark.parametrize("n_fft", [None, 512])
def test_frozen_global_statistics_with_recipe_specaug_vs_oracle(n_fft, mode):
    """The ASR recipes' training step once the global statistics are frozen (normalize: True = global statistics
    averaged over max_epoch_num epochs; their SpecAugment block leaves time_warp on): the frontend kernel writes every
    row normalised and frequency-masked, a small kernel warps and time-masks rows [0, minlen) in place."""
    g = torch.Generator().manual_seed(777)
    lens = sorted(torch.randint(16000, 64001, (7,), generator=g).tolist(), reverse=True)
    wav = _wave(23, lens, zero_seg=False)
    conf = dict(n_mels=80, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    if n_fft:
        conf["n_fft"] = n_fft
    sconf = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5, time_warp_mode=mode)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="global", max_epoch_num=1), specaug=sconf).to(DEV)
    chain.train()
    x, xl = wav.to(DEV)[..., None], torch.tensor(lens, device=DEV)
    torch.manual_seed(5)
    chain(x, xl, epoch=1)                       # one averaging step fixes the statistics
    torch.manual_seed(99)
    y, yl = chain(x, xl, epoch=2)               # epoch > max_epoch_num: statistics known up front
    feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
    ref = O.feat_norm_global([(feat, flen)], O.RunningStats(), 1)[0]   # the statistics of that one step
    d = O.specaug_draw(O.TorchCpuRandint(99), len(lens), 80, flen, O.SpecAugCfg(**sconf))
    assert d.warp is not None
    ref = O.specaug_apply(ref, flen, d, feat_norm=True)
    got = y.cpu().numpy()
    assert np.array_equal(yl.cpu().numpy(), flen)
    mask = O.specaug_mask(*got.shape, d)
    assert (got[mask] == 0.0).all() and np.count_nonzero((got == 0.0) & ~mask) <= 2
    assert_close(got, ref, what=f"frozen global statistics + warp ({mode}, n_fft {n_fft})")


@pytest.mark.parametrize("stats", ["both", "mean_only", "std_only"])
def test_averaging_epochs_through_the_chain_vs_oracle(stats):
    """The ASR recipes' training step during the first max_epoch_num epochs (normalize: True = global statistics, still
    averaged; SpecAugment with time warp): frontend launch with sums + packed batch statistics, running average with
    the counter, normalise + warp + masks — against the oracle step by step, buffers and counter included, then the
    state dict into a fresh module (the counter must continue) and a step with epoch=None (count restarts at 1)."""
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    sconf = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
    nconf = dict(norm_type="global", mean_norm=stats != "std_only", std_norm=stats != "mean_only")
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=nconf, specaug=sconf).to(DEV)
    chain.train()
    st = O.RunningStats()
    g = torch.Generator().manual_seed(4321)

    def step(chain, k, epoch):
        lens = sorted(torch.randint(16000, 56001, (6,), generator=g).tolist(), reverse=True)
        wav = _wave(60 + k, lens, zero_seg=False)
        torch.manual_seed(100 + k)
        y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV), epoch=epoch)
        feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
        ms, ss = O.utterance_stats(feat, flen, 1e-10)
        bm, bs = O.batch_stats([ms], [ss])
        st.register(bm, bs, epoch)
        ref = feat
        if nconf["mean_norm"]:
            ref = ref - st.mean
        if nconf["std_norm"]:
            ref = ref / st.std
        d = O.specaug_draw(O.TorchCpuRandint(100 + k), len(lens), 80, flen, O.SpecAugCfg(**sconf))
        assert d.warp is not None
        ref = O.specaug_apply(ref, flen, d, feat_norm=True)
        got = y.cpu().numpy()
        assert np.array_equal(yl.cpu().numpy(), flen)
        mask = O.specaug_mask(*got.shape, d)
        assert (got[mask] == 0.0).all()
        assert_close(got, ref, what=f"averaging step {k} ({stats})")
        n = chain.normalize
        if nconf["mean_norm"]:
            assert_close(n.global_mean.cpu().numpy(), st.mean, what=f"global_mean {k}")
        else:
            assert not hasattr(n, "global_mean")
        if nconf["std_norm"]:
            assert_close(n.global_std.cpu().numpy(), st.std, what=f"global_std {k}")
        else:
            assert not hasattr(n, "global_std")
        return n

    for k in range(3):
        n = step(chain, k, epoch=1 + k)
        assert int(n.global_batch.item()) == k + 1
    fresh = AcousticFrontend(frontend=dict(conf=conf), normalize=nconf, specaug=sconf).to(DEV)
    fresh.load_state_dict(chain.state_dict())
    fresh.train()
    n = step(fresh, 3, epoch=4)
    assert int(n.global_batch.item()) == 4
    n = step(fresh, 4, epoch=None)  # feat_norm.py:469-472: no epoch -> the counter restarts, the average is the batch
    assert int(n.global_batch.item()) == 1


@pytest.mark.parametrize("warp", [False, True])
def test_batch_statistics_in_training_through_the_chain_vs_oracle(warp):
    """norm_type='batch' in training (feat_norm.py:232-299): the batch's own statistics, nothing kept between steps."""
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    sconf = dict(time_warp=warp, freq_mask_width=20, freq_mask_num=2, time_mask_width=0.05, time_mask_num=3)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="batch"), specaug=sconf).to(DEV)
    chain.train()
    g = torch.Generator().manual_seed(99)
    for k in range(2):
        lens = sorted(torch.randint(12000, 50001, (5,), generator=g).tolist(), reverse=True)
        wav = _wave(80 + k, lens, zero_seg=False)
        torch.manual_seed(40 + k)
        y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV), epoch=1)
        feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
        ref = O.feat_norm_batch([(feat, flen)], training=True)[0]
        d = O.specaug_draw(O.TorchCpuRandint(40 + k), len(lens), 80, flen, O.SpecAugCfg(**sconf))
        assert (d.warp is not None) == warp
        ref = O.specaug_apply(ref, flen, d, feat_norm=True)
        assert np.array_equal(yl.cpu().numpy(), flen)
        assert_close(y.cpu().numpy(), ref, what=f"batch statistics, step {k}, warp {warp}")
    assert not [n for n, _ in chain.normalize.named_buffers()]  # nothing is remembered


def test_c5_style_linear_and_mel_from_one_launch_with_running_global_norm():
    # BASELINE config 5 in miniature: magnitude linear spectrogram AND mel from one launch; mel -> global norm with
    # the running average over several batches (epoch=1) -> SpecAugment
    conf = dict(sr=22050, n_mels=80, win_length=1024, hop_length=256, mag_spec=True, fmin=0, fmax=8000, log_base=None,
                clamp=1e-5)
    fe = Speech2MelSpec(**conf).to(DEV)
    gn = FeatureNormalization(input_size=80, norm_type="global").to(DEV)
    gn.train()
    sa = SpecAugment(input_size=80, time_warp=False).to(DEV)
    st = O.RunningStats()
    g = torch.Generator().manual_seed(555)
    torch.manual_seed(7)
    rng = O.TorchCpuRandint(7)
    for step in range(3):
        lens = sorted(torch.randint(2 * 22050, 5 * 22050, (5,), generator=g).tolist(), reverse=True)
        wav = _wave(30 + step, lens, zero_seg=False)
        out = fe.run(wav.to(DEV), torch.tensor(lens, device=DEV), want_lin=True, want_stats=True)
        y = out["mel"]
        y._sfe_stats = (y._version, out["stats"])
        _, masks = sa.draw_device(len(lens), 80, out["feat_len_host"], DEV)
        y, _ = gn.forward(y, out["feat_len"], epoch=1, masks=masks, inplace=True)
        cfg = O.FrontendCfg(**conf)
        lin_ref, flen, _ = O.speech2linear(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**{**cfg.__dict__}))
        assert_close(out["lin"].cpu().numpy(), lin_ref, what=f"linear {step}")
        feat, _, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, cfg)
        ref = O.feat_norm_global([(feat, flen)], st, 1)[0]
        d = O.specaug_draw(rng, len(lens), 80, flen, O.SpecAugCfg(time_warp=False))
        ref = O.specaug_apply(ref, flen, d, feat_norm=True)
        assert_close(y.cpu().numpy(), ref, what=f"mel {step}")
        assert int(gn.global_batch.item()) == step + 1


def test_checkpoint_load_rebuilds_the_plan():
    # load_state_dict copies into the filterbank buffer in place: the cached plan must not go stale
    g = load_golden("fe_c1")
    m = Speech2MelSpec(**FE_CASES["fe_c1"]).to(DEV)
    wav, wl = cu(g["wav"])[..., None], cu(g["wav_len"])
    a, _ = m(wav, wl)
    sd = m.state_dict()
    sd["linear2mel.mel_fbanks.weight"] = sd["linear2mel.mel_fbanks.weight"] * 2.0
    m.load_state_dict(sd, strict=True)
    b, _ = m(wav, wl)
    valid = int(g["feat_len"][0])
    assert torch.allclose(b[0, :valid][a[0, :valid] > -9.9], a[0, :valid][a[0, :valid] > -9.9] + np.log10(2.0), atol=1e-4)


def test_pcm16_upload_equals_the_float_path():
    """16-bit PCM staged at two bytes per sample and converted on the GPU must be bit-identical to what soundfile's
    float32 read (x / 32768) gives the reference's loader — and to the fp32 upload of those floats."""
    g = torch.Generator().manual_seed(21)
    lens = [30001, 22222, 9999, 17, 1]
    pcm = torch.randint(-32768, 32768, (len(lens), max(lens), 1), generator=g, dtype=torch.int32).to(torch.int16)
    pcm[0, :8, 0] = torch.tensor([-32768, 32767, 0, 1, -1, 12345, -12345, 256], dtype=torch.int16)
    wl = torch.tensor(lens, dtype=torch.int64)
    dev = torch.full(pcm.shape, float("nan"), device=DEV)
    ops.upload_waveforms_i16(pcm.pin_memory(), wl, out=dev)
    want = pcm.float() / 32768.0
    for i, n in enumerate(lens):
        assert torch.equal(dev[i, :n].cpu(), want[i, :n])
        assert torch.isnan(dev[i, n:]).all()  # padding is neither transferred nor written
    m = Speech2MelSpec(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, preemphasis=0.97).to(DEV)
    a, _ = m(dev, wl.to(DEV))
    b, _ = m(ops.upload_waveforms(want.pin_memory(), wl, device=DEV), wl.to(DEV))
    assert torch.equal(a, b)
    # an odd pitch (rows not 16-byte aligned on the device) takes the scalar path
    odd = torch.full((len(lens), max(lens) + 3), float("nan"), device=DEV)[:, 1:1 + max(lens)]
    assert ops.upload_i16_stage_bytes(len(lens), max(lens)) > 2 * sum(lens)


@pytest.mark.parametrize("norm", [None, dict(norm_type="utterance"), dict(norm_type="global")])
def test_bf16_output_is_the_fp32_output_rounded(norm):
    """the chain can hand the prenet bfloat16 (prenet/conv2d.py:220-276 runs under autocast) — it
    must be exactly the fp32 result rounded to nearest even, masks included, in training and in eval mode."""
    lens = [48000, 40001, 33333, 8000]
    wav = _wave(31, lens).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    aug = dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)
    a = AcousticFrontend(frontend=dict(conf=conf), normalize=norm, specaug=aug if norm is not None else None).to(DEV)
    b = AcousticFrontend(frontend=dict(conf=conf), normalize=norm, specaug=aug if norm is not None else None,
                         out_dtype="bfloat16").to(DEV)
    for mode in ("train", "eval"):
        getattr(a, mode)()
        getattr(b, mode)()
        torch.manual_seed(5)
        y32, l32 = a(wav[..., None], wl, epoch=None)
        torch.manual_seed(5)
        y16, l16 = b(wav[..., None], wl, epoch=None)
        assert y16.dtype == torch.bfloat16 and y16.shape == y32.shape and torch.equal(l16, l32)
        assert torch.equal(y16, y32.to(torch.bfloat16)), f"{norm} {mode}"


def test_an_older_larger_plan_still_launches_after_a_smaller_one_was_created():
    """cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel FUNCTION, and plans share kernel instances
    (same n_fft): a later, smaller plan must not shrink the limit under an earlier one (recover() builds a no-mel
    plan next to the training plan; two frontends may differ only in hop or n_mels)."""
    lens = [24000, 16000]
    wav = _wave(33, lens).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    big = Speech2MelSpec(n_mels=80, n_fft=512, win_length=400, hop_length=160, preemphasis=0.97).to(DEV)
    a, _ = big(wav[..., None], wl)
    # smaller users of the same kernel functions: shorter hop (smaller staging), fewer mels, no mel stage at all
    for small in (Speech2MelSpec(n_mels=80, n_fft=512, win_length=400, hop_length=80, preemphasis=0.97),
                  Speech2MelSpec(n_mels=40, n_fft=512, win_length=400, hop_length=160, preemphasis=0.97),
                  Speech2LinearSpec(n_fft=512, win_length=400, hop_length=160, preemphasis=0.97)):
        small.to(DEV)(wav[..., None], wl)
        b, _ = big(wav[..., None], wl)
        torch.cuda.synchronize()
        assert torch.equal(a, b)


@pytest.mark.parametrize("skip_pad", [False, True])
def test_host_made_work_split_equals_the_kernels_own(skip_pad):
    """`wav_len_host` only moves the cutting of the batch into per-SM ranges from the kernel's prologue to the host:
    features, feature lengths and the fixed-point sums must be bit-identical with and without it (ragged batch with
    an utterance shorter than a frame hop, an empty one, and more utterances than a warp scans at once)."""
    g = torch.Generator().manual_seed(77)
    lens = sorted(torch.randint(3000, 90000, (70,), generator=g).tolist(), reverse=True)
    lens[0], lens[-1], lens[-2] = 90000, 0, 100
    wav = _wave(51, lens, zero_seg=False).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    host = np.asarray(lens, dtype=np.int64)
    m = Speech2MelSpec(n_mels=80, n_fft=512, win_length=400, hop_length=160, preemphasis=0.97).to(DEV)
    plan = m.speech2linear._plan(wav.device, preemph=True, mel=m.linear2mel.mel_conf())
    F = plan.num_frames(wav.size(1))
    outs = []
    for lh in (None, host):
        mel = torch.full((len(lens), F, 80), -7.0, device=DEV)  # rows the kernel skips keep the marker in both runs
        o = plan.forward(wav, wl, want_stats=True, skip_pad=skip_pad, out_mel=mel, len_host=lh)
        torch.cuda.synchronize()
        outs.append(o)
    a, b = outs
    assert torch.equal(a["feat_len"], b["feat_len"])
    assert torch.equal(a["stats"], b["stats"])
    assert torch.equal(a["mel"], b["mel"])
    # and the split really was the host's: lengths that disagree with the device's change what is computed
    # (not under the A/B switches that take the host-made split or the tile kernel out of the picture)
    import os
    if os.environ.get("SFE_HOST_SPLIT") == "0" or os.environ.get("SFE_KERNEL", "").startswith("g"):
        return
    wrong = host.copy()
    wrong[0] = 45000
    mel = torch.full((len(lens), F, 80), -7.0, device=DEV)
    c = plan.forward(wav, wl, want_stats=True, skip_pad=skip_pad, out_mel=mel, len_host=wrong)
    torch.cuda.synchronize()
    assert not torch.equal(c["mel"][0], outs[1]["mel"][0])


def test_rows_past_each_length_are_written_by_the_frontend_kernel_itself():
    """Ragged batch on the recipe configuration WITHOUT a normalise pass behind the kernel (the module-by-module
    drop-in): the rows past each utterance's frames — log(clamp) for the Mel output, 0 for the energy — are written by
    the kernel's FFT warps once they have run out of pairs.  The outputs come from `torch.empty`: poison the allocator's
    free blocks first so that a row nobody writes cannot pass as a zero."""
    g = torch.Generator().manual_seed(321)
    lens = sorted(torch.randint(200, 48000, (37,), generator=g).tolist(), reverse=True)
    lens[0], lens[-1] = 48000, 0
    wav = _wave(61, lens, zero_seg=False)
    conf = dict(n_mels=80, n_fft=512, win_length=400, hop_length=160, sr=16000, return_energy=True)
    cfg = O.FrontendCfg(**conf)
    ref, rl, ren = O.speech2mel(wav.numpy().astype(np.float64), lens, cfg)
    m = Speech2MelSpec(**conf).to(DEV)
    wl = torch.tensor(lens, device=DEV)
    x = wav.to(DEV)[..., None]
    for _ in range(2):
        poison = [torch.full((n,), float("nan"), device=DEV) for n in (37 * 301 * 80, 37 * 301, 37 * 301 * 80)]
        torch.cuda.synchronize()
        del poison
        feat, flen, en, _ = m(x, wl)
        torch.cuda.synchronize()
        assert np.array_equal(flen.cpu().numpy(), rl)
        assert not torch.isnan(feat).any() and not torch.isnan(en).any()
        assert_close(feat.cpu().numpy(), ref, what="mel incl. padded rows")
        assert_close(en.cpu().numpy(), ren, what="energy incl. padded rows")


def test_recipe_specaug_on_a_batch_beyond_the_inline_descriptor_block():
    """2 * B * (4 + 5) mask descriptors of the recipes' SpecAugment block fit the kernel parameter block for up to 128
    utterances; a larger batch uploads them (DeviceMasks, device arrays) — same masks, same features."""
    B = 150
    g = torch.Generator().manual_seed(150)
    lens = sorted(torch.randint(3000, 12000, (B,), generator=g).tolist(), reverse=True)
    wav = _wave(77, lens, zero_seg=False)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    sconf = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"), specaug=sconf).to(DEV)
    chain.train()
    torch.manual_seed(31)
    y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
    ref = O.feat_norm_utterance(feat, flen)
    d = O.specaug_draw(O.TorchCpuRandint(31), B, 80, flen, O.SpecAugCfg(**sconf))
    ref = O.specaug_apply(ref, flen, d, feat_norm=True)
    got = y.cpu().numpy()
    mask = O.specaug_mask(*got.shape, d)
    assert (got[mask] == 0.0).all() and np.count_nonzero((got == 0.0) & ~mask) <= 2
    assert_close(got, ref, what="recipe SpecAugment, B = 150")


def test_time_warp_over_more_rows_than_shared_memory_stages():
    """The in-place warp stages rows [0, minlen) of a CTA's Mel bins in shared memory; where minlen rows do not fit
    (here: > 12 800 frames, two utterances of over two minutes) the chain falls back to one out-of-place gather pass
    over the whole tensor.  Same result either way."""
    lens = [2_100_000, 2_090_000]
    wav = _wave(91, lens, zero_seg=False)
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    sconf = dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"), specaug=sconf).to(DEV)
    chain.train()
    torch.manual_seed(17)
    y, yl = chain(wav.to(DEV)[..., None], torch.tensor(lens, device=DEV))
    feat, flen, _ = O.speech2mel(wav.numpy().astype(np.float64), lens, O.FrontendCfg(**conf))
    assert int(flen.min()) * 16 > 200 * 1024
    ref = O.feat_norm_utterance(feat, flen)
    d = O.specaug_draw(O.TorchCpuRandint(17), len(lens), 80, flen, O.SpecAugCfg(**sconf))
    assert d.warp is not None
    ref = O.specaug_apply(ref, flen, d, feat_norm=True)
    got = y.cpu().numpy()
    mask = O.specaug_mask(*got.shape, d)
    assert (got[mask] == 0.0).all()
    assert_close(got, ref, what="warp fallback (long utterances)")


def test_fresh_cuda_length_tensor_every_step_never_sees_the_previous_batch():
    """What the real caller does (model/abs.py:329-352): a NEW CUDA length tensor with other contents every step, the
    old one freed in between — so that the allocator hands out the same block and CPython the same id().  The host
    copy of the lengths rides on the tensor object and dies with it: mask bounds, frame counts and the host-made work
    split of every step must be those of THAT step's lengths (checked against the oracle step by step)."""
    conf = dict(n_mels=80, n_fft=512, win_length=0.025, hop_length=0.010, sr=16000, preemphasis=0.97)
    sconf = dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2)
    chain = AcousticFrontend(frontend=dict(conf=conf), normalize=dict(norm_type="utterance"), specaug=sconf).to(DEV)
    chain.train()
    cfg = O.FrontendCfg(**conf)
    g = torch.Generator().manual_seed(5150)
    T = 40000
    wav = 0.1 * torch.randn(6, T, generator=g)
    x = wav.to(DEV)[..., None]
    ptrs = set()
    for step in range(6):
        lens = torch.randint(4000, T + 1, (6,), generator=g).tolist()
        lens[step % 6] = T
        xl = torch.tensor(lens, dtype=torch.int64).to(DEV)   # fresh device tensor, no host copy remembered
        ptrs.add(xl.data_ptr())
        torch.manual_seed(900 + step)
        y, yl = chain(x, xl)
        w = wav.clone().numpy().astype(np.float64)
        for i, n in enumerate(lens):
            w[i, n:] = 0.0
        feat, flen, _ = O.speech2mel(w, lens, cfg)
        ref = O.feat_norm_utterance(feat, flen)
        d = O.specaug_draw(O.TorchCpuRandint(900 + step), 6, 80, flen, O.SpecAugCfg(**sconf))
        ref = O.specaug_apply(ref, flen, d, feat_norm=True)
        got = y.cpu().numpy()
        assert np.array_equal(yl.cpu().numpy(), flen), f"step {step}: frame counts"
        mask = O.specaug_mask(*got.shape, d)
        assert (got[mask] == 0.0).all() and np.count_nonzero((got == 0.0) & ~mask) <= 2, f"step {step}: mask pattern"
        assert_close(got, ref, what=f"step {step}")
        del xl, y, yl
    assert len(ptrs) < 6, "the allocator was expected to reuse the length tensor's block (the stale-cache scenario)"

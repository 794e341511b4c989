"""Host-side mirror of the reference module API: constructor keys, derived sizes, state-dict keys,
reprs, filterbank bit-equality, the SpecAugment draw sequence — everything that runs without a GPU."""
import copy
import os
import re

import numpy as np
import pytest
import torch

from oracle import frontend_np as O
from sfe_testutil import load_golden
from speechain_b200.module.acoustic_frontend import AcousticFrontend
from speechain_b200.module.augment.specaug import SpecAugment
from speechain_b200.module.frontend._common import as_float, make_window, mel_filterbank, to_samples
from speechain_b200.module.frontend.delta_feat import DeltaFeature
from speechain_b200.module.frontend.linear2mel import LinearSpec2MelSpec
from speechain_b200.module.frontend.speech2linear import Speech2LinearSpec
from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
from speechain_b200.module.norm.feat_norm import FeatureNormalization
from test_oracle_golden import FE_CASES, SA_CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_seconds_vs_samples_and_nfft_default():
    assert to_samples(0.025, 16000) == 400 and to_samples(0.010, 16000) == 160
    assert to_samples(0.05, 22050) == 1102 and to_samples(256, 22050) == 256
    m = Speech2MelSpec(n_mels=80, hop_length=0.010, win_length=0.025, preemphasis=0.97)
    assert m.speech2linear.n_fft == 400 and m.speech2linear.output_size == 201 and m.output_size == 80
    assert Speech2MelSpec(n_mels=80, hop_length=160, win_length=400, delta_order=2).output_size == 240
    assert m.get_sample_rate() == 16000


def test_yaml_string_numbers_are_coerced():
    # PyYAML reads `clamp: 1e-5` as a str (the reference's ruamel loader as a float)
    assert as_float("1e-5", "clamp") == 1e-5
    m = Speech2MelSpec(n_mels=80, hop_length=256, win_length=1024, sr=22050, clamp="1e-5", log_base=None)
    assert m.linear2mel.clamp == 1e-5 and m.linear2mel.log_base is None
    with pytest.raises(ValueError):
        as_float("abc", "clamp")


@pytest.mark.parametrize("name", sorted(FE_CASES))
def test_filterbank_is_bit_identical_to_the_reference_buffer(name):
    g = load_golden(name)
    c = O.FrontendCfg(**FE_CASES[name])
    fmax = c.fmax if c.fmax is not None else c.sr // 2
    W = mel_filterbank(c.sr, c.n_fft, c.n_mels, float(c.fmin), float(fmax), c.mel_scale, c.mel_norm)
    assert np.array_equal(W.numpy(), g["mel_weight"])


def test_filterbank_matches_torchaudio_when_present():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(sample_rate=16000, n_mels=80, n_freqs=201, f_min=0.0, f_max=8000,
                                        norm="slaney", mel_scale="slaney").T
    assert torch.equal(mel_filterbank(16000, 400, 80, 0.0, 8000.0, "slaney", True), ref)


def test_window_table():
    assert torch.equal(make_window("hann", 400, 512, False), torch.hann_window(400))
    assert torch.allclose(make_window("hamming", 400, 512, True), torch.hamming_window(400) * 512 ** -0.5)
    with pytest.raises(ValueError):
        make_window("nosuch", 400, 512, False)


def test_state_dict_keys_and_reprs_match_the_reference():
    m = Speech2MelSpec(n_mels=80, hop_length=0.010, win_length=0.025, preemphasis=0.97)
    sd = m.state_dict()
    assert sorted(sd) == ["linear2mel.mel_fbanks.weight", "speech2linear.preemph_filter.weight"]
    assert tuple(sd["speech2linear.preemph_filter.weight"].shape) == (1, 1, 2)
    assert torch.allclose(sd["speech2linear.preemph_filter.weight"].flatten(), torch.tensor([-0.97, 1.0]))
    assert tuple(sd["linear2mel.mel_fbanks.weight"].shape) == (80, 201)
    assert sorted(Speech2MelSpec(n_mels=80, hop_length=160, win_length=400).state_dict()) == ["linear2mel.mel_fbanks.weight"]
    assert list(m.parameters()) == []  # nothing trainable
    r = repr(m)
    assert "Speech2MelSpec(" in r and "win_length=400, hop_length=160, n_fft=400, mag_spec=False, preemphasis=0.97" in r
    assert "stft_dim=201, n_mels=80, fmin=0.0, fmax=8000, mel_scale=slaney, mel_norm=True" in r
    d = DeltaFeature(delta_order=2)
    assert list(d.state_dict()) == ["delta_filters.weight"] and tuple(d.delta_filters.weight.shape) == (1, 1, 5, 1)
    assert "norm_type=global, mean_norm=True, std_norm=True" in repr(FeatureNormalization(input_size=80))
    assert "freq_mask_width=[0, 15]" in repr(SpecAugment(input_size=80, freq_mask_width=15))


def test_constructor_errors_mirror_the_reference():
    with pytest.raises(AssertionError):
        LinearSpec2MelSpec(n_fft=512, n_mels=80, fmin=9000)              # fmax > fmin (linear2mel.py:122-124)
    with pytest.raises(AssertionError):
        LinearSpec2MelSpec(n_fft=512, n_mels=80, mel_scale="bark")       # linear2mel.py:127-130
    with pytest.raises(AssertionError):
        SpecAugment(input_size=80, freq_mask_width=80)                   # specaug.py:53-57
    with pytest.raises(AssertionError):
        SpecAugment(time_warp=False, freq_mask=False, time_mask=False)   # specaug.py:34-36
    with pytest.raises(ValueError):
        SpecAugment(input_size=80, freq_mask_width="wide")
    with pytest.raises(AssertionError):
        DeltaFeature(delta_order=3)
    with pytest.raises(ValueError):
        Speech2LinearSpec(hop_length=160, win_length=400, pre_stft_norm="robust")


def test_cpu_tensors_fail_loudly():
    m = Speech2MelSpec(n_mels=80, hop_length=160, win_length=400, n_fft=512)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(2, 4000, 1), torch.tensor([4000, 4000]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        FeatureNormalization(input_size=80, norm_type="utterance")(torch.zeros(2, 10, 80), torch.tensor([10, 8]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        SpecAugment(input_size=80)(torch.zeros(2, 50, 80), torch.tensor([50, 40]))
    with pytest.raises(RuntimeError, match="multi-channel"):
        m(torch.zeros(2, 4000, 2), torch.tensor([4000, 4000]))


@pytest.mark.parametrize("name", sorted(SA_CASES))
def test_specaug_draw_sequence_equals_oracle(name):
    g = load_golden(name)
    conf = SA_CASES[name]
    sa = SpecAugment(input_size=80, **conf)
    B, _, D = g["x"].shape
    torch.manual_seed(int(g["seed"]))
    warp, fpos, flen, tpos, tlen = sa.draw(B, D, int(g["x_len"].min()))
    d = O.specaug_draw(O.TorchCpuRandint(int(g["seed"])), B, D, g["x_len"], O.SpecAugCfg(**conf))
    assert (warp is None) == (d.warp is None)
    if warp is not None:
        assert tuple(warp[:2]) == tuple(d.warp)
    for a, b in ((fpos, d.freq_pos), (flen, d.freq_len), (tpos, d.time_pos), (tlen, d.time_len)):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)


def test_modules_survive_deepcopy_and_chain_builds_from_yaml_blocks():
    chain = AcousticFrontend(
        frontend=dict(type="frontend.speech2mel.Speech2MelSpec",
                      conf=dict(sr=16000, preemphasis=0.97, hop_length=0.010, win_length=0.025, n_mels=80)),
        normalize=True, specaug=dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5))
    assert chain.normalize.norm_type == "global" and chain.specaug.feat_norm and chain.output_size == 80
    c2 = copy.deepcopy(chain)
    assert sorted(c2.state_dict()) == sorted(chain.state_dict())


def test_product_code_never_touches_the_oracle():
    # oracle/ is test infrastructure: nothing under speechain_b200/ may import or reference it
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "speechain_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "ref_port_torch" in src:
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_upload_row_groups_bound_the_padding():
    """sfe_upload_bytes mirrors the grouping of sfe_upload_waveforms (no GPU needed): equal rows cost exactly their
    samples, ragged rows at most 24/23 of them (a group's padding stays below 1/24 of its bytes), isolated rows nothing
    extra."""
    import torch

    from speechain_b200.host import ops
    assert ops.upload_bytes(torch.tensor([100, 100, 100])) == 4 * 300
    assert ops.upload_bytes(torch.tensor([1000, 10])) == 4 * 1010          # too different: two copies
    assert ops.upload_bytes(torch.tensor([0, 0])) == 0
    g = torch.Generator().manual_seed(3)
    lens = torch.tensor(sorted(torch.randint(32000, 320001, (64,), generator=g).tolist(), reverse=True))
    valid = 4 * int(lens.sum())
    assert valid <= ops.upload_bytes(lens) <= valid * 24 // 23 + 4


@pytest.mark.parametrize("helper", [False, True])
def test_lean_draw_path_equals_reference_draw_sequence(helper, monkeypatch):
    """draw_device's persistent-buffer path consumes the CPU generator exactly like ``draw`` (the reference's call
    sequence): same warp, same intervals, step after step, for several configurations — with the tensor-op draws
    (``Tensor.random_`` in place) and with the one-call helper (csrc/sfe_draw.cpp) where it is built."""
    import numpy as np
    import torch

    from speechain_b200.module.augment import specaug as SA
    from speechain_b200.module.augment.specaug import SpecAugment
    if helper:
        monkeypatch.setattr(SA, "_DRAW", [False, None])
        if SA._draw_helper() is None:
            pytest.skip("draw helper not built")
    else:
        monkeypatch.setattr(SA, "_DRAW", [True, None])
    for conf in (dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2),
                 dict(time_warp=True, freq_mask_width=15, freq_mask_num=4, time_mask_width=20, time_mask_num=4),
                 dict(time_warp=True, freq_mask=False, time_mask_width=[2, 0.1], time_mask_num=3),
                 dict(time_warp=False, time_mask=False, freq_mask_width=[3, 10], freq_mask_num=1)):
        a, b = SpecAugment(input_size=80, **conf), SpecAugment(input_size=80, **conf)
        rng = np.random.default_rng(1)
        torch.manual_seed(11)
        ref = []
        lens = [rng.integers(40, 900, size=16) for _ in range(4)]
        for fl in lens:
            ref.append(a.draw(16, 80, int(fl.min())))
        torch.manual_seed(11)
        for fl, (warp, fpos, flen, tpos, tlen) in zip(lens, ref):
            w2, m = b.draw_device(16, 80, fl, "cpu")
            assert w2 == warp and m.on_host
            got = {"fpos": m.views[0], "flen": m.views[1], "tpos": m.views[2], "tlen": m.views[3]}
            for name, want in (("fpos", fpos), ("flen", flen), ("tpos", tpos), ("tlen", tlen)):
                if want is None:
                    assert got[name] is None
                else:
                    assert np.array_equal(got[name].numpy(), want), name
            st = m.struct()
            assert st.n_freq == (0 if fpos is None else fpos.shape[1]) and st.n_time == (0 if tpos is None else tpos.shape[1])
            assert st.on_host == 1 and st.fill == 0.0


def test_host_length_cache_dies_with_its_tensor():
    """A length tensor built fresh every step (model/abs.py:329-352) may reuse the id(), the allocator block and
    the version counter of the previous step's tensor: the cached host lengths must not survive that."""
    from speechain_b200.host.ops import HostLengths

    class FakeCuda:  # the attributes HostLengths looks at, without a device
        is_cuda, _version = True, 0

        def __init__(self, values):
            self.values = values

        def data_ptr(self):
            return 0x7F0000000000  # the caching allocator hands the same block back

        def cpu(self):
            return torch.tensor(self.values)

    a = FakeCuda([5, 4, 3])
    HostLengths.remember(a, np.array([5, 4, 3]))
    assert HostLengths.get(a).tolist() == [5, 4, 3]
    del a
    b = FakeCuda([9, 8, 7])  # may well sit at the same id(); same data_ptr and version
    assert HostLengths.get(b).tolist() == [9, 8, 7]
    # an in-place edit or a re-pointed tensor invalidates the copy
    b._version = 1
    b.values = [1, 1, 1]
    assert HostLengths.get(b).tolist() == [1, 1, 1]
    # real tensors: the copy rides on the object
    t = torch.tensor([3, 2, 1])
    assert HostLengths.get(t).tolist() == [3, 2, 1]


def test_draw_helper_makes_the_reference_draw_sequence():
    """csrc/sfe_draw.cpp draws every mask of a step in one call on torch's default CPU generator: same values, same
    generator state afterwards as the reference's torch.randint sequence (specaug.py:98-108,145-158,178-191) —
    for the recipe block (time warp on), with and without each kind of mask, float and integer widths."""
    import copy
    from speechain_b200.csrc import build as B
    from speechain_b200.module.augment import specaug as SA
    try:
        B.build_draw_ext()
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"draw helper cannot be built here: {e}")
    SA._DRAW[0], SA._DRAW[1] = False, None
    assert SA._draw_helper() is not None
    confs = [dict(freq_mask_width=15, freq_mask_num=4, time_mask_width=0.05, time_mask_num=5),
             dict(time_warp=False, freq_mask_width=30, freq_mask_num=2, time_mask_width=0.05, time_mask_num=2),
             dict(time_warp=True, freq_mask=False, time_mask_width=[3, 20], time_mask_num=3, time_warp_window=8),
             dict(time_warp=False, time_mask=False, freq_mask_width=[2, 9], freq_mask_num=1),
             dict(time_warp=True, freq_mask=False, time_mask=False)]
    rng = np.random.default_rng(3)
    for conf in confs:
        a = SA.SpecAugment(input_size=80, **conf)
        b = copy.deepcopy(a)  # (before the first draw: no persistent arrays yet)
        for step in range(4):
            fl = rng.integers(12 if step == 3 else 40, 400, size=16)
            torch.manual_seed(100 + step)
            w1, fpos, flen, tpos, tlen = a.draw(16, 80, int(fl.min()))
            s1 = torch.get_rng_state()
            torch.manual_seed(100 + step)
            w2, m = b.draw_device(16, 80, fl, "cpu")
            s2 = torch.get_rng_state()
            assert torch.equal(s1, s2), "the generator must be left where the reference leaves it"
            assert (w1 is None) == (w2 is None) and (w1 is None or tuple(w1[:3]) == tuple(w2[:3]))
            got = iter(m._host) if m is not None else iter(())
            for ref in (fpos, flen, tpos, tlen):
                if ref is not None:
                    assert np.array_equal(next(got), ref)
    # a pickled / deep-copied module starts with fresh descriptor arrays (they hold raw host addresses)
    c = copy.deepcopy(b)
    assert "_mask_cache" not in c.__dict__


def test_which_training_steps_average_and_which_are_frozen():
    """Host-side schedule choice of FeatureNormalization (feat_norm.py:455-508): a training step updates the running
    statistics while epoch <= max_epoch_num (or no epoch is given); afterwards the statistics are frozen — but only
    once all the buffers the step needs exist, otherwise the reference's own path (and its error) is taken."""
    n = FeatureNormalization(input_size=4, norm_type="global", max_epoch_num=2)
    n.train()
    assert n.averaging("global", None) and n.averaging("global", 1) and n.averaging("global", 2)
    assert not n.averaging("global", 3) and not n.averaging(None, 1)
    assert not n.frozen("global", 3)                       # nothing registered yet
    n.register_buffer("global_mean", torch.zeros(4))
    assert not n.averaging("global", 1)                    # one of the two buffers only: the general path
    n.register_buffer("global_std", torch.ones(4))
    assert n.averaging("global", 1)
    assert not n.frozen("global", 3)                       # no counter yet
    n.register_buffer("global_batch", torch.tensor([7]))
    assert n.frozen("global", 3) and not n.frozen("global", 2) and not n.frozen("global", None)
    n.eval()
    assert not n.averaging("global", 1)
    b = FeatureNormalization(input_size=4, norm_type="batch")
    b.train()
    assert not b.averaging(None, 1) and not b.frozen(None, 9)
    m = FeatureNormalization(input_size=4, norm_type="group", std_norm=False, max_epoch_num=2)
    m.train()
    assert m.averaging("spk3", 1)
    m.register_buffer("spk3_mean", torch.zeros(4))
    m.register_buffer("spk3_batch", torch.tensor([1]))
    assert m.averaging("spk3", 2) and m.frozen("spk3", 3) and not m.frozen("spk4", 3)
    # the host mirror of the counter: restored from the buffer once, then advanced without a sync
    assert n._counter("global") == 7

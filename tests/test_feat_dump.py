"""Offline feature dumper (SURVEY §8 f4 iii): host logic on the CPU, end-to-end parity on the GPU."""
import wave

import numpy as np
import pytest

from speechain_b200.host import feat_dump as FD


def _write_wav(path, x, sr, width=2):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(width)
        w.setframerate(sr)
        if width == 2:
            w.writeframes((np.clip(x, -1, 1 - 1 / 32768) * 32768).astype("<i2").tobytes())
        else:
            w.writeframes((np.clip(x, -1, 1 - 2.0 ** -31) * 2147483648.0).astype("<i4").tobytes())


def test_idx_files_and_wav_reader(tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1234) * 0.1).astype(np.float32)
    p16, p32 = str(tmp_path / "a.wav"), str(tmp_path / "b b.wav")
    _write_wav(p16, x, 16000, 2)
    _write_wav(p32, x, 22050, 4)
    y, sr = FD.read_wav(p16)
    assert sr == 16000 and y.dtype == np.float32 and np.abs(y - np.round(x * 32768) / 32768).max() <= 1 / 32768
    y, sr = FD.read_wav(p32)
    assert sr == 22050 and np.abs(y - x).max() < 1e-6
    idx = tmp_path / "idx2wav"
    idx.write_text(f"utt1 {p16}\nutt0 {p32}\n")          # a path with a blank: split at the FIRST separator
    d = FD.load_idx2data_file(str(idx))
    assert d == {"utt1": p16, "utt0": p32}
    out = FD.write_idx2feat(str(tmp_path), [("utt1", "/x/utt1.npz"), ("utt0", "/x/utt0.npz")])
    assert open(out).read().splitlines() == ["utt0 /x/utt0.npz", "utt1 /x/utt1.npz"]
    i, f = FD.save_feat(str(tmp_path), "utt9", np.ones((3, 2), np.float32), 16000)
    z = np.load(f)
    assert i == "utt9" and z["feat"].shape == (3, 2) and int(z["sample_rate"]) == 16000


def test_groups_hold_one_length_each():
    lens = [500, 300, 500, 500, 7, 300]
    groups = FD.group_by_length(lens, max_batch=2)
    assert sorted(i for g in groups for i in g) == list(range(len(lens)))
    for g in groups:
        assert len(g) <= 2 and len({lens[i] for i in g}) == 1
    assert [lens[g[0]] for g in groups] == sorted([lens[g[0]] for g in groups], reverse=True)


@pytest.mark.gpu
def test_dump_equals_solo_module_runs(tmp_path):
    import torch

    from speechain_b200.module.frontend.speech2mel import Speech2MelSpec
    rng = np.random.default_rng(1)
    lens = [8000, 6500, 8000, 4321, 6500, 8000]
    lines = []
    for i, n in enumerate(lens):
        p = str(tmp_path / f"u{i}.wav")
        _write_wav(p, (rng.standard_normal(n) * 0.1).astype(np.float32), 16000)
        lines.append(f"u{i} {p}")
    (tmp_path / "idx2wav").write_text("\n".join(lines) + "\n")
    conf = dict(n_mels=80, hop_length=0.010, win_length=0.025, preemphasis=0.97, sr=8000)  # sr is overridden
    # (chunk=4: the index is walked in two chunks — host memory is bounded by a chunk, not by the corpus)
    out = FD.dump_features(str(tmp_path / "idx2wav"), "logmel", conf, str(tmp_path / "feat"), chunk=4)
    rows = [r.split(" ", 1) for r in open(out).read().splitlines()]
    assert [r[0] for r in rows] == sorted(f"u{i}" for i in range(len(lens)))
    m = Speech2MelSpec(n_mels=80, hop_length=0.010, win_length=0.025, preemphasis=0.97, sr=16000).to("cuda:0")
    for idx, path in rows:
        z = np.load(path)
        wav, sr = FD.read_wav(str(tmp_path / f"{idx}.wav"))
        assert int(z["sample_rate"]) == sr == 16000
        ref, rl = m(torch.from_numpy(wav)[None, :, None].to("cuda:0"), torch.tensor([len(wav)], device="cuda:0"))
        assert z["feat"].shape == (int(rl[0]), 80)
        assert np.array_equal(z["feat"], ref[0].cpu().numpy()), "a dumped utterance must equal its solo run bit for bit"
    # a second call finds idx2feat and skips
    assert FD.dump_features(str(tmp_path / "idx2wav"), "logmel", conf, str(tmp_path / "feat")) is None

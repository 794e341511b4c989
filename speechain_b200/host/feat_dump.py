"""Offline feature extraction on the GPU: the B200 replacement of the reference's CPU-pool dumper
(datasets/pyscripts/feat_extractor.py:42-84).

Same inputs and outputs: an ``idx2wav`` text file (``<index> <path>`` per line, load_idx2data_file:
speechain/utilbox/data_loading_util.py:91-140), one ``<feat_path>/<index>.npz`` per utterance holding ``feat``
(frames, dim) and ``sample_rate`` (feat_extractor.py:62), and a sorted ``idx2feat`` index file (:81); an existing
``idx2feat`` skips the run (:70).  The features themselves come from this package's frontend modules, i.e. they are
the ones the models compute on the fly (``Speech2MelSpec`` / ``Speech2LinearSpec`` semantics), not librosa's.

Every utterance must come out exactly as if it had been processed alone (the reference extracts one file at a
time).  The frontend reflects the *batch* tensor at its right edge (speech2linear.py:198), so only utterances as
long as the batch behave like a solo run: utterances are therefore grouped by identical length and every group
is one launch — many short launches, each a few tens of microseconds; the disk is the bottleneck, which a thread
pool of writers hides.

Host logic (index files, grouping, WAV decoding, npz writing) has no GPU dependency and is tested on the CPU;
``dump_features`` needs a CUDA device like everything else in this package.
"""
import os
import wave
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def load_idx2data_file(path: str, separator: str = " ") -> Dict[str, str]:
    """``<index><separator><value>`` per line, split at the first separator (data_loading_util.py:120-140)."""
    assert os.path.exists(path), f"{path} doesn't exist!"
    out = {}
    with open(path, mode="r") as f:
        for row in f:
            row = row.replace("\n", "")
            if not row:
                continue
            idx, val = row.split(separator, 1)
            out[idx] = val
    return out


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """PCM WAV -> (samples (T,) float32 in [-1, 1), sample rate); first channel of multi-channel files.
    (The reference reads through soundfile with dtype float32, data_loading_util.py:71-80: same scaling.)"""
    if path.lower().endswith(".npy"):
        raise ValueError("read_wav expects a .wav file")
    with wave.open(path, "rb") as w:
        sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float32) / 8388608.0
    else:
        raise NotImplementedError(f"{width * 8}-bit PCM")
    if nch > 1:
        x = x.reshape(-1, nch)[:, 0]
    return np.ascontiguousarray(x), sr


def group_by_length(lengths: Sequence[int], max_batch: int = 256) -> List[List[int]]:
    """Positions grouped so that every group holds utterances of ONE length (see the module docstring),
    at most ``max_batch`` per group, longest first."""
    by_len = defaultdict(list)
    for i, n in enumerate(lengths):
        by_len[int(n)].append(i)
    groups = []
    for n in sorted(by_len, reverse=True):
        ids = by_len[n]
        for s in range(0, len(ids), max_batch):
            groups.append(ids[s:s + max_batch])
    return groups


def save_feat(feat_path: str, index: str, feat: np.ndarray, sample_rate: int) -> Tuple[str, str]:
    """One ``<index>.npz`` with ``feat`` and ``sample_rate`` (feat_extractor.py:62-63)."""
    out = os.path.join(feat_path, f"{index}.npz")
    np.savez(out, feat=feat, sample_rate=sample_rate)
    return index, out


def write_idx2feat(feat_path: str, pairs: Iterable[Tuple[str, str]]) -> str:
    """Sorted ``<index> <npz path>`` lines, exactly what np.savetxt(fmt='%s') writes (feat_extractor.py:81)."""
    out = os.path.join(feat_path, "idx2feat")
    np.savetxt(out, sorted(pairs, key=lambda x: x[0]), fmt="%s")
    return out


def build_module(feat_type: str, feat_config: dict, sample_rate: int):
    """feat_type 'logmel' -> Speech2MelSpec, 'stft' -> Speech2LinearSpec (feat_extractor.py:49-56; 'mfcc' has no
    counterpart among the frontend modules).  ``sr`` of the config is replaced by the file's own rate (:45-46)."""
    conf = dict(feat_config)
    conf.pop("sr", None)
    if feat_type == "logmel":
        from ..module.frontend.speech2mel import Speech2MelSpec
        return Speech2MelSpec(sr=sample_rate, **conf)
    if feat_type == "stft":
        from ..module.frontend.speech2linear import Speech2LinearSpec
        return Speech2LinearSpec(sr=sample_rate, **conf)
    raise ValueError(f"Unknown acoustic feature: {feat_type}! It should be one of ['stft', 'logmel'].")


def dump_features(idx2wav: str, feat_type: str, feat_config: dict, feat_path: str, device="cuda:0", max_batch: int = 256,
                  writers: int = 8, reader: Optional[Callable[[str], Tuple[np.ndarray, int]]] = None,
                  chunk: int = 2048) -> Optional[str]:
    """Extracts ``feat_type`` features of every utterance of ``idx2wav`` on ``device`` and writes the reference's
    file layout under ``feat_path``.  Returns the path of ``idx2feat`` (None if it already existed: run skipped).

    The index is walked in chunks of ``chunk`` files (decoded by the thread pool, grouped by sampling rate and
    identical length, launched, written): host memory is bounded by one chunk of waveforms plus the features still
    on their way to disk, whatever the size of the corpus (the reference extracts file by file)."""
    import torch

    os.makedirs(feat_path, exist_ok=True)
    if os.path.exists(os.path.join(feat_path, "idx2feat")):
        print(f"Acoustic features have already existed in {feat_path}, so the extraction process is skipped.")
        return None
    reader = reader or read_wav
    items = list(load_idx2data_file(idx2wav).items())
    modules, pairs = {}, []
    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        for c0 in range(0, len(items), max(1, chunk)):
            part = items[c0:c0 + max(1, chunk)]
            loaded = list(pool.map(lambda kv: reader(kv[1]), part))
            by_sr = defaultdict(list)
            for i, (_, sr) in enumerate(loaded):
                by_sr[int(sr)].append(i)
            pending = []
            for sr, ids in by_sr.items():
                module = modules.get(sr)
                if module is None:
                    module = modules[sr] = build_module(feat_type, feat_config, sr).to(device).eval()
                for grp in group_by_length([len(loaded[i][0]) for i in ids], max_batch):
                    rows = [ids[j] for j in grp]
                    wav = torch.from_numpy(np.stack([loaded[i][0] for i in rows])).to(device)[..., None]
                    n = wav.size(1)
                    if n == 0:
                        raise ValueError(f"empty waveform: {part[rows[0]][0]}")
                    out = module(wav, torch.full((len(rows),), n, dtype=torch.int64, device=device))
                    feat, flen = out[0].cpu().numpy(), out[1].cpu().numpy()
                    for r, i in enumerate(rows):
                        pending.append(pool.submit(save_feat, feat_path, part[i][0], feat[r, :int(flen[r])].copy(), sr))
            pairs += [f.result() for f in pending]  # (the chunk's files are on disk before the next one is decoded)
    return write_idx2feat(feat_path, pairs)


def main(argv=None):
    """Command line of the reference script (feat_extractor.py:27-39); --ncpu becomes the number of writer threads."""
    import argparse

    import yaml
    ap = argparse.ArgumentParser(description="GPU acoustic feature dumper")
    ap.add_argument("--idx2wav", type=str, required=True)
    ap.add_argument("--feat_type", type=str, required=True)
    ap.add_argument("--feat_config", type=str, required=True)
    ap.add_argument("--feat_path", type=str, required=True)
    ap.add_argument("--ncpu", type=int, default=8)
    ap.add_argument("--device", type=str, default="cuda:0")
    a = ap.parse_args(argv)
    with open(a.feat_config) as f:
        conf = yaml.safe_load(f)
    dump_features(a.idx2wav, a.feat_type, conf, a.feat_path, device=a.device, writers=a.ncpu)


if __name__ == "__main__":
    main()

"""CPU oracle for the acoustic-frontend hot path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing in ``speechain_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

This is an *explicit* restatement of what the reference modules compute, written as flat
numpy functions (no torch in the arithmetic), selectable fp32 / fp64:

* ``speech2linear``   <- speechain/module/frontend/speech2linear.py:142-250
* ``mel_filterbank``  <- torchaudio.functional.melscale_fbanks (third-party, torchaudio
                         2.11.0 in this image; called at linear2mel.py:135-143)
* ``linear2mel``      <- speechain/module/frontend/linear2mel.py:163-189
* ``speech2mel``      <- speechain/module/frontend/speech2mel.py:158-189
* ``delta_feature``   <- speechain/module/frontend/delta_feat.py:57-81
* ``utterance_stats`` / ``feat_norm_*`` <- speechain/module/norm/feat_norm.py:104-336,455-508
* ``TorchCpuRandint`` <- torch.randint on the CPU mt19937 generator (third-party, torch
                         2.11.0; call sites specaug.py:98-108,145-158,178-191)
* ``specaug_draw`` / ``specaug_apply`` <- speechain/module/augment/specaug.py:86-206
* ``time_warp``       <- specaug.py:89-127 + torch ``upsample_bicubic2d`` (A=-0.75,
                         align_corners=True)

``torch.stft`` (third-party; call site speech2linear.py:198) is restated as: batch-level
reflect padding by n_fft//2, frames of n_fft samples every ``hop`` samples, the window
zero-padded centrally to n_fft, one-sided DFT, no normalisation.

Parity pin: the reference holds no tests or golden vectors for this path.
The oracle is pinned against outputs of the reference modules themselves, run in the build
container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against those fixtures.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# configuration helpers
# --------------------------------------------------------------------------------------


def to_samples(x, sr: int) -> int:
    """float => seconds, int => samples (speech2linear.py:85-90)."""
    return int(x * sr) if isinstance(x, float) else int(x)


@dataclass
class FrontendCfg:
    """Resolved constructor arguments of Speech2LinearSpec / Speech2MelSpec."""

    hop_length: int
    win_length: int
    n_fft: Optional[int] = None
    sr: int = 16000
    preemphasis: Optional[float] = None
    pre_stft_norm: Optional[str] = None
    window: str = "hann"
    center: bool = True
    normalized: bool = False
    onesided: bool = True
    mag_spec: bool = False
    return_energy: bool = False
    # linear-stage log (Speech2LinearSpec standalone)
    lin_clamp: float = 1e-10
    lin_logging: bool = False
    lin_log_base: Optional[float] = None
    # mel stage
    n_mels: Optional[int] = None
    fmin: float = 0.0
    fmax: Optional[float] = None
    clamp: float = 1e-10
    logging: bool = True
    log_base: Optional[float] = 10.0
    mel_scale: str = "slaney"
    mel_norm: bool = True
    delta_order: Optional[int] = None
    delta_N: int = 2

    def __post_init__(self):
        self.hop_length = to_samples(self.hop_length, self.sr)
        self.win_length = to_samples(self.win_length, self.sr)
        if self.n_fft is None:
            self.n_fft = self.win_length  # speech2linear.py:93-94


# --------------------------------------------------------------------------------------
# waveform stage
# --------------------------------------------------------------------------------------


def preemphasis(wav: np.ndarray, lens: np.ndarray, coeff: float) -> np.ndarray:
    """y[0]=x[0], y[n]=x[n]-a*x[n-1], then y[i, len_i:]=0 (speech2linear.py:167-173)."""
    a = wav.dtype.type(coeff)
    y = wav.copy()
    y[:, 1:] = wav[:, 1:] - a * wav[:, :-1]
    for i, n in enumerate(lens):
        y[i, int(n):] = 0
    return y


def pre_stft_norm(wav: np.ndarray, kind: str) -> np.ndarray:
    """speech2linear.py:177-187.  Statistics include the padding.  ``mean_std`` only
    broadcasts for B == 1 in the reference (missing keepdim); same restriction here."""
    if kind == "min_max":
        lo = wav.min(axis=1, keepdims=True)
        hi = wav.max(axis=1, keepdims=True)
        return (wav - lo) / (hi - lo) * 2 - 1
    if kind == "mean_std":
        if wav.shape[0] != 1:
            raise RuntimeError("pre_stft_norm='mean_std' only broadcasts for batch size 1")
        return (wav - wav.mean(axis=1, keepdims=True)) / wav.std(axis=1, ddof=1, keepdims=True)
    raise ValueError(kind)


def periodic_window(name: str, win: int, dtype) -> np.ndarray:
    n = np.arange(win, dtype=np.float64)
    if name == "hann":
        w = 0.5 - 0.5 * np.cos(2 * np.pi * n / win)
    elif name == "hamming":
        w = 0.54 - 0.46 * np.cos(2 * np.pi * n / win)
    elif name == "blackman":
        w = 0.42 - 0.5 * np.cos(2 * np.pi * n / win) + 0.08 * np.cos(4 * np.pi * n / win)
    elif name == "bartlett":
        w = 1.0 - np.abs(2.0 * n / win - 1.0)
    else:
        raise ValueError(f"unsupported window {name!r}")
    return w.astype(dtype)


def num_frames(T: int, n_fft: int, hop: int, center: bool) -> int:
    """Frames torch.stft produces for a length-T row."""
    if center:
        T = T + 2 * (n_fft // 2)
    return 1 + (T - n_fft) // hop


def feat_lengths(lens: np.ndarray, n_fft: int, hop: int, center: bool) -> np.ndarray:
    """speech2linear.py:201-212 (without mutating the caller's lengths)."""
    lens = np.asarray(lens, dtype=np.int64)
    if center:
        lens = lens + 2 * (n_fft // 2)
    return (lens - n_fft) // hop + 1


def stft_power(wav: np.ndarray, cfg: FrontendCfg) -> np.ndarray:
    """|STFT|^2 of a zero-padded batch (B, T) -> (B, F, n_fft//2+1).

    Restates torch.stft(center, pad_mode='reflect', onesided=True): the reflect padding is
    applied to the *batch tensor*, so only rows as long as the batch see their own samples
    mirrored at the right edge."""
    B, T = wav.shape
    n_fft, hop, win = cfg.n_fft, cfg.hop_length, cfg.win_length
    dt = wav.dtype
    w = np.zeros(n_fft, dtype=dt)
    left = (n_fft - win) // 2
    w[left:left + win] = periodic_window(cfg.window, win, dt)
    if cfg.normalized:
        w = (w * dt.type(n_fft ** -0.5)).astype(dt)
    if cfg.center:
        pad = n_fft // 2
        wav = np.pad(wav, ((0, 0), (pad, pad)), mode="reflect")
    F = 1 + (wav.shape[1] - n_fft) // hop
    out = np.empty((B, F, n_fft // 2 + 1), dtype=dt)
    for b in range(B):
        fr = np.lib.stride_tricks.sliding_window_view(wav[b], n_fft)[::hop][:F]
        spec = np.fft.rfft(fr * w, axis=-1)
        out[b] = (spec.real.astype(dt)) ** 2 + (spec.imag.astype(dt)) ** 2
    return out


def speech2linear(wav: np.ndarray, lens: Sequence[int], cfg: FrontendCfg):
    """Returns (linear_spec (B,F,K), feat_len (B,), energy (B,F) or None)."""
    lens = np.asarray(lens, dtype=np.int64)
    if wav.ndim == 3:
        if wav.shape[-1] != 1:
            raise RuntimeError("multi-channel speech is not supported")
        wav = wav[..., 0]
    if cfg.preemphasis is not None:
        wav = preemphasis(wav, lens, cfg.preemphasis)
    if cfg.pre_stft_norm is not None:
        wav = pre_stft_norm(wav, cfg.pre_stft_norm)
    spec = stft_power(wav, cfg)
    if not cfg.onesided:
        # torch.stft(onesided=False) returns all n_fft bins; for a real signal X[n_fft - k] = conj(X[k]), so the power
        # of bins K .. n_fft-1 mirrors bins K-2 .. 1 (odd n_fft: K-1 .. 1).  speech2linear.py:197-198
        K = spec.shape[-1]
        spec = np.concatenate([spec, spec[..., cfg.n_fft - K:0:-1]], axis=-1)
    flen = feat_lengths(lens, cfg.n_fft, cfg.hop_length, cfg.center)
    energy = None
    if cfg.return_energy:
        energy = np.sqrt(np.maximum(spec.sum(axis=-1), spec.dtype.type(1e-10)))
    for i, n in enumerate(flen):
        spec[i, int(n):] = 0
        if energy is not None:
            energy[i, int(n):] = 0
    if cfg.mag_spec:
        spec = np.sqrt(spec)
    if cfg.lin_logging:
        spec = np.log(np.maximum(spec, spec.dtype.type(cfg.lin_clamp)))
        if cfg.lin_log_base is not None:
            spec = spec / spec.dtype.type(math.log(cfg.lin_log_base))
    return spec, flen, energy


# --------------------------------------------------------------------------------------
# mel stage
# --------------------------------------------------------------------------------------


def _hz_to_mel(f: float, scale: str) -> float:
    if scale == "htk":
        return 2595.0 * math.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mel = f / f_sp
    if f >= 1000.0:
        mel = 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return mel


def _mel_to_hz(m: np.ndarray, scale: str) -> np.ndarray:
    if scale == "htk":
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    f = f_sp * m
    min_log_mel = 1000.0 / f_sp
    logstep = math.log(6.4) / 27.0
    hi = m >= min_log_mel
    f = np.where(hi, 1000.0 * np.exp(logstep * (m - min_log_mel)), f)
    return f


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0,
                   fmax: Optional[float] = None, mel_scale: str = "slaney",
                   mel_norm: bool = True, dtype=np.float64) -> np.ndarray:
    """(n_mels, n_fft//2+1) weights; published algorithm of torchaudio's melscale_fbanks
    (linspace 0..sr//2, mel points, max(0, min(down, up)) triangles, Slaney area norm)."""
    K = n_fft // 2 + 1
    fmax = float(sr // 2) if fmax is None else float(fmax)
    freqs = np.linspace(0.0, float(sr // 2), K)
    m_pts = np.linspace(_hz_to_mel(float(fmin), mel_scale), _hz_to_mel(fmax, mel_scale), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, mel_scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    if mel_norm:
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb.T.astype(dtype)


def linear2mel(spec: np.ndarray, weight: np.ndarray, clamp: float = 1e-10, logging: bool = True,
               log_base: Optional[float] = 10.0) -> np.ndarray:
    """spec (B,F,K) @ weight.T, clamp, log, /ln(base) (linear2mel.py:177-187)."""
    dt = spec.dtype
    mel = spec @ weight.T.astype(dt)
    if logging:
        mel = np.log(np.maximum(mel, dt.type(clamp)))
        if log_base is not None:
            mel = mel / dt.type(math.log(log_base))
    return mel


def mel2linear(mel: np.ndarray, flen: Sequence[int], weight: np.ndarray, logging: bool = True,
               log_base: Optional[float] = 10.0, clamp: float = 1e-10) -> np.ndarray:
    """LinearSpec2MelSpec.recover (linear2mel.py:191-224): undo the log, minimum-norm least squares
    ``W x = mel`` per frame (what lstsq(driver="gels") returns for the full-row-rank n_mels x K filterbank),
    silence rows >= len of every utterance shorter than the longest, clamp at 1e-10.  fp64."""
    x = mel.astype(np.float64)
    if logging:
        x = np.power(np.e if log_base is None else float(log_base), x)
    W = weight.astype(np.float64)
    pinv = W.T @ np.linalg.inv(W @ W.T)  # (K, D)
    lin = x @ pinv.T
    flen = np.asarray(flen)
    for i, n in enumerate(flen):
        if n != flen.max():
            lin[i, n:] = 0.0
    return np.maximum(lin, clamp)


def griffinlim(spec: np.ndarray, n_fft: int, hop: int, win: int, window: np.ndarray, power: float,
               angles0: Optional[np.ndarray], n_iter: int = 32, momentum: float = 0.99) -> np.ndarray:
    """torchaudio.functional.griffinlim (torchaudio 2.x; the reference calls it through
    transforms.GriffinLim at speech2linear.py:281-293) restated with numpy FFTs in fp64.

    spec (B, F, K) power (power=2) or magnitude (power=1) spectrogram; angles0 (B, F, K) complex initial
    estimate (torch.rand: NOT unit length) or None for ones.  Returns (B, hop*(F-1)).
    torch.istft: irfft of every frame, times the (centrally zero-padded) window, overlap-add, divide by the
    overlap-added squared window, trim n_fft/2 per side.  torch.stft: reflect-pad n_fft/2, frame, window, rfft."""
    B, F, K = spec.shape
    mag = spec.astype(np.float64) ** (1.0 / power)
    mom = momentum / (1.0 + momentum)
    wpad = np.zeros(n_fft)
    lp = (n_fft - win) // 2
    wpad[lp:lp + win] = window.astype(np.float64)
    L = hop * (F - 1)

    def istft(S):
        S = S.copy()
        S[..., 0] = S[..., 0].real  # C2R transforms ignore the imaginary parts of DC and Nyquist
        S[..., -1] = S[..., -1].real
        fr = np.fft.irfft(S, n=n_fft, axis=-1) * wpad
        out = np.zeros((B, n_fft + L))
        env = np.zeros(n_fft + L)
        for t in range(F):
            out[:, t * hop:t * hop + n_fft] += fr[:, t]
            env[t * hop:t * hop + n_fft] += wpad ** 2
        return out[:, n_fft // 2:n_fft // 2 + L] / env[n_fft // 2:n_fft // 2 + L]

    def stft(x):
        xp = np.pad(x, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
        idx = np.arange(n_fft)[None, :] + hop * np.arange(F)[:, None]
        return np.fft.rfft(xp[:, idx] * wpad, axis=-1)

    angles = np.ones((B, F, K), np.complex128) if angles0 is None else angles0.astype(np.complex128)
    tprev = 0.0
    for _ in range(n_iter):
        rebuilt = stft(istft(mag * angles))
        angles = rebuilt - mom * tprev
        angles = angles / (np.abs(angles) + 1e-16)
        tprev = rebuilt
    return istft(mag * angles)


def speech2linear_recover(spec: np.ndarray, flen: Sequence[int], n_fft: int, hop: int, win: int, window: np.ndarray,
                          mag_spec: bool, preemphasis: Optional[float], angles0: Optional[np.ndarray],
                          inv_preemph_winlen: int = 100):
    """Speech2LinearSpec.recover (speech2linear.py:252-351): Griffin-Lim, inverse pre-emphasis over a sliding
    window of `inv_preemph_winlen` samples (y[n] = sum_j a^j x[n-j]), silence beyond (len-1)*hop."""
    wav = griffinlim(spec, n_fft, hop, win, window, 1.0 if mag_spec else 2.0, angles0)
    wav_len = (np.asarray(flen) - 1) * hop
    if preemphasis is not None:
        taps = np.float32(preemphasis) ** np.arange(inv_preemph_winlen, dtype=np.float32)  # fp32 like torch.pow
        y = np.zeros_like(wav)
        for j, a in enumerate(taps.astype(np.float64)):
            y[:, j:] += a * wav[:, :wav.shape[1] - j]
        wav = y
    for i, n in enumerate(wav_len):
        wav[i, n:] = 0.0
    return wav, wav_len


def delta_feature(feat: np.ndarray, order: int, N: int = 2) -> np.ndarray:
    """Delta / delta-delta by a (2N+1)-tap FIR along time with zero padding, concatenated
    on the feature axis (delta_feat.py:34-78)."""
    dt = feat.dtype
    taps = (np.arange(-N, N + 1) / (2 * sum(i * i for i in range(1, N + 1)))).astype(dt)

    def fir(x):
        B, F, D = x.shape
        xp = np.zeros((B, F + 2 * N, D), dtype=dt)
        xp[:, N:N + F] = x
        y = np.zeros_like(x)
        for j in range(2 * N + 1):  # cross-correlation, like Conv2d
            y += taps[j] * xp[:, j:j + F]
        return y

    stack = [feat, fir(feat)]
    if order == 2:
        stack.append(fir(stack[1]))
    return np.concatenate(stack, axis=-1)


def speech2mel(wav: np.ndarray, lens: Sequence[int], cfg: FrontendCfg, weight: Optional[np.ndarray] = None):
    """Returns (feat (B,F,D), feat_len, energy or None).  ``weight`` lets the caller pass
    the reference's own fp32 filterbank so that only the arithmetic is compared."""
    lin_cfg = FrontendCfg(**{**cfg.__dict__, "lin_logging": False})  # speech2mel.py:136
    spec, flen, energy = speech2linear(wav, lens, lin_cfg)
    if weight is None:
        weight = mel_filterbank(cfg.sr, cfg.n_fft, cfg.n_mels, cfg.fmin, cfg.fmax, cfg.mel_scale,
                                cfg.mel_norm, dtype=spec.dtype)
    feat = linear2mel(spec, weight, cfg.clamp, cfg.logging, cfg.log_base)
    if cfg.delta_order is not None:
        feat = delta_feature(feat, cfg.delta_order, cfg.delta_N)
    return feat, flen, energy


# --------------------------------------------------------------------------------------
# feature normalisation
# --------------------------------------------------------------------------------------


def utterance_stats(feat: np.ndarray, flen: Sequence[int], clamp: float = 1e-10):
    """Per-utterance mean and unbiased std over valid frames (feat_norm.py:115-133)."""
    means = np.stack([feat[i, :int(n)].mean(axis=0) for i, n in enumerate(flen)])
    with np.errstate(invalid="ignore", divide="ignore"):
        stds = np.stack([feat[i, :int(n)].std(axis=0, ddof=1) for i, n in enumerate(flen)])
    stds = np.maximum(stds, feat.dtype.type(clamp))  # NaN (1 valid frame) propagates like torch.clamp
    return means, stds


def feat_norm_utterance(feat, flen, mean_norm=True, std_norm=True, clamp=1e-10):
    """feat_norm.py:137-139: applied to padded frames too."""
    means, stds = utterance_stats(feat, flen, clamp)
    out = feat
    if mean_norm:
        out = out - means[:, None, :]
    if std_norm:
        out = out / stds[:, None, :]
    return out


@dataclass
class RunningStats:
    """The lazily registered ``{prefix}_batch/_mean/_std`` buffers (feat_norm.py:455-508)."""

    max_epoch_num: int = 4
    batch: Optional[int] = None
    mean: Optional[np.ndarray] = None
    std: Optional[np.ndarray] = None

    def register(self, cur_mean, cur_std, epoch):
        if epoch is None or self.batch is None:
            self.batch = 1
        elif epoch <= self.max_epoch_num:
            self.batch += 1
        if epoch is None or epoch <= self.max_epoch_num:
            # torch: 1 / LongTensor -> float32 tensor
            w = np.float32(1.0) / np.float32(self.batch)
            if cur_mean is not None:
                w_ = cur_mean.dtype.type(w)
                self.mean = cur_mean if self.mean is None else w_ * cur_mean + (1 - w_) * self.mean
            if cur_std is not None:
                w_ = cur_std.dtype.type(w)
                self.std = cur_std if self.std is None else w_ * cur_std + (1 - w_) * self.std


def batch_stats(shard_means: List[np.ndarray], shard_stds: List[np.ndarray]):
    """Mean of per-utterance stats over all ranks: sum_r(sum_i mu_i) / sum_r N_r
    (feat_norm.py:237-275); with one shard this is ``curr_means.mean(0)``."""
    n = sum(m.shape[0] for m in shard_means)
    dt = shard_means[0].dtype
    if len(shard_means) == 1:
        return shard_means[0].mean(axis=0), shard_stds[0].mean(axis=0)
    msum = np.stack([m.sum(axis=0) for m in shard_means]).sum(axis=0) / dt.type(n)
    ssum = np.stack([s.sum(axis=0) for s in shard_stds]).sum(axis=0) / dt.type(n)
    return msum, ssum


def feat_norm_global(shards: List[Tuple[np.ndarray, Sequence[int]]], state: RunningStats, epoch,
                     training=True, clamp=1e-10):
    """``norm_type='global'`` over W emulated ranks (one (feat, flen) per rank).
    Training: stats of all ranks -> running average -> normalise by the updated buffers
    (feat_norm.py:233-336).  Eval: normalise by the stored buffers."""
    if training:
        ms, ss = zip(*[utterance_stats(f, l, clamp) for f, l in shards])
        bm, bs = batch_stats(list(ms), list(ss))
        state.register(bm, bs, epoch)
    return [(f - state.mean) / state.std for f, _ in shards]


def feat_norm_batch(shards, training=True, clamp=1e-10):
    """``norm_type='batch'``: training -> batch stats, eval -> utterance stats
    (feat_norm.py:283-299)."""
    if not training:
        return [feat_norm_utterance(f, l, clamp=clamp) for f, l in shards]
    ms, ss = zip(*[utterance_stats(f, l, clamp) for f, l in shards])
    bm, bs = batch_stats(list(ms), list(ss))
    return [(f - bm) / bs for f, _ in shards]


# --------------------------------------------------------------------------------------
# SpecAugment
# --------------------------------------------------------------------------------------


class TorchCpuRandint:
    """``torch.randint`` on torch's default CPU generator, restated.

    torch.manual_seed(s) seeds an mt19937 with the classic ``init_genrand(s)``; every
    int64 draw with ``high - low < 2**32`` consumes one 32-bit output ``r`` and returns
    ``r % (high - low) + low`` in C order."""

    def __init__(self, seed: int):
        self._bg = np.random.MT19937()
        self._bg._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def randint(self, low: int, high: int, size) -> np.ndarray:
        n = int(np.prod(size))
        raw = self._bg.random_raw(n).astype(np.uint64)
        rng = np.uint64(high - low)
        return (raw % rng).astype(np.int64).reshape(size) + low


@dataclass
class SpecAugCfg:
    time_warp: bool = True
    time_warp_window: int = 5
    time_warp_mode: str = "bicubic"
    freq_mask: bool = True
    freq_mask_width: object = 30
    freq_mask_num: int = 2
    time_mask: bool = True
    time_mask_width: object = 0.05
    time_mask_num: int = 2
    feat_norm: bool = True

    def __post_init__(self):
        if isinstance(self.freq_mask_width, int):
            self.freq_mask_width = [0, self.freq_mask_width]
        if isinstance(self.time_mask_width, (int, float)):
            self.time_mask_width = [0, self.time_mask_width]


@dataclass
class SpecAugDraw:
    warp: Optional[Tuple[int, int]] = None          # (warp_center, warp_pos)
    freq_pos: Optional[np.ndarray] = None           # (B, Nf)
    freq_len: Optional[np.ndarray] = None
    time_pos: Optional[np.ndarray] = None           # (B, Nt)
    time_len: Optional[np.ndarray] = None
    time_minlen: int = 0
    warp_mode: str = "bicubic"


def specaug_draw(rng: TorchCpuRandint, B: int, D: int, flen: Sequence[int], cfg: SpecAugCfg) -> SpecAugDraw:
    """The exact randint call sequence of specaug.py:96-108,145-158,167-191."""
    minlen = int(min(flen))
    d = SpecAugDraw(time_minlen=minlen, warp_mode=cfg.time_warp_mode)
    if cfg.time_warp and minlen > 2 * cfg.time_warp_window + 1:
        w = cfg.time_warp_window
        center = int(rng.randint(w + 1, minlen - w, (1,))[0])
        pos = int(rng.randint(center - w, center + w, (1,))[0])
        d.warp = (center, pos)
    if cfg.freq_mask:
        lo, hi = cfg.freq_mask_width
        d.freq_len = rng.randint(lo, hi + 1, (B, cfg.freq_mask_num))
        d.freq_pos = rng.randint(0, max(1, D - int(d.freq_len.max())), (B, cfg.freq_mask_num))
    if cfg.time_mask:
        lo, hi = cfg.time_mask_width
        if isinstance(lo, float):
            lo = int(lo * minlen)
        if isinstance(hi, float):
            hi = int(hi * minlen)
        hi = min(hi, minlen)
        d.time_len = rng.randint(lo, hi + 1, (B, cfg.time_mask_num))
        d.time_pos = rng.randint(0, max(1, minlen - int(d.time_len.max())), (B, cfg.time_mask_num))
    return d


def _cubic_coeffs(t: np.ndarray, A: float = -0.75):
    """Cubic-convolution weights of torch's upsample_bicubic2d for offsets -1,0,1,2."""
    def c1(x):  # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1
    def c2(x):  # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)


def _resample_time(x: np.ndarray, out_len: int, mode: str = "bicubic") -> np.ndarray:
    """(B, L, D) -> (B, out_len, D): bicubic (torch upsample_bicubic2d) or bilinear (upsample_bilinear2d: taps i0 and
    min(i0 + 1, L - 1) with weights 1 - t, t) with align_corners=True along time only; the feature axis keeps its
    size so its interpolation is the identity."""
    B, L, D = x.shape
    dt = x.dtype
    if out_len == 0:
        return np.zeros((B, 0, D), dtype=dt)
    scale = dt.type((L - 1) / (out_len - 1)) if out_len > 1 else dt.type(0)
    src = (scale * np.arange(out_len, dtype=dt)).astype(dt)
    i0 = np.floor(src).astype(np.int64)
    t = (src - i0.astype(dt)).astype(dt)
    if mode == "bilinear":
        i1 = np.minimum(i0 + 1, L - 1)
        return ((dt.type(1) - t)[None, :, None] * x[:, i0, :] + t[None, :, None] * x[:, i1, :]).astype(dt)
    assert mode == "bicubic", mode
    ws = _cubic_coeffs(t)
    out = np.zeros((B, out_len, D), dtype=dt)
    for k, w in enumerate(ws):
        idx = np.clip(i0 - 1 + k, 0, L - 1)
        out += w.astype(dt)[None, :, None] * x[:, idx, :]
    return out


def time_warp(feat: np.ndarray, center: int, pos: int, minlen: int, mode: str = "bicubic") -> np.ndarray:
    """specaug.py:111-124: [0,center)->[0,pos) and [center,minlen)->[pos,minlen)."""
    out = feat.copy()
    out[:, :pos] = _resample_time(feat[:, :center], pos, mode)
    out[:, pos:minlen] = _resample_time(feat[:, center:minlen], minlen - pos, mode)
    return out


def specaug_mask(B: int, F: int, D: int, d: SpecAugDraw) -> np.ndarray:
    """Boolean (B,F,D) mask: union of frequency stripes (all frames, padded ones too) and
    time stripes (specaug.py:160-163,193-199)."""
    mask = np.zeros((B, F, D), dtype=bool)
    if d.freq_pos is not None:
        k = np.arange(D)[None, None, :]
        fm = (d.freq_pos[:, :, None] <= k) & (k < (d.freq_pos + d.freq_len)[:, :, None])
        mask |= fm.any(axis=1)[:, None, :]
    if d.time_pos is not None:
        t = np.arange(F)[None, :, None]
        tm = (d.time_pos[:, None, :] <= t) & (t < (d.time_pos + d.time_len)[:, None, :])
        mask |= tm.any(axis=-1)[:, :, None]
    return mask


def specaug_apply(feat: np.ndarray, flen: Sequence[int], d: SpecAugDraw, feat_norm: bool = True) -> np.ndarray:
    B, F, D = feat.shape
    if d.warp is not None:
        feat = time_warp(feat, d.warp[0], d.warp[1], d.time_minlen, d.warp_mode)
    mask = specaug_mask(B, F, D, d)
    if d.freq_pos is None and d.time_pos is None:
        return feat
    fill = feat.dtype.type(0.0) if feat_norm else feat.mean(dtype=feat.dtype)
    return np.where(mask, fill, feat)

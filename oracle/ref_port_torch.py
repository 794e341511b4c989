"""CPU baseline port: the reference frontend's op sequence on torch CPU tensors.

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing in ``speechain_b200/`` may import this.

``/root/reference`` does not travel to the GPU box, so ``bench.py``'s ``cpu_baseline`` and
``--impl reference`` legs time this port instead (``kind: "port"``).  It issues the same
third-party calls, in the same order and with the same Python per-utterance loops, as the
reference modules do on CPU — ``torch.stft`` (speech2linear.py:198), the dense
``feat @ W.T`` (linear2mel.py:177), per-utterance ``mean``/``std`` slices
(feat_norm.py:115-133), ``torch.randint`` + ``masked_fill`` (specaug.py:145-204) — so its
wall-clock is what the reference costs on the same cores, and (checked by
``tests/golden/make_golden.py`` in the build container) its outputs are bit-identical to
the reference modules'.  It is written as flat functions over an explicit state dict, not
as nn.Modules.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def mel_weight(sr, n_fft, n_mels, fmin=0.0, fmax=None, mel_scale="slaney", mel_norm=True) -> torch.Tensor:
    """(n_mels, n_fft//2+1) fp32 filterbank with torchaudio's op order (linear2mel.py:135-143)."""
    K = n_fft // 2 + 1
    fmax = sr // 2 if fmax is None else fmax

    def hz2mel(f):
        if mel_scale == "htk":
            return 2595.0 * math.log10(1.0 + (f / 700.0))
        f_sp = 200.0 / 3
        mels = (f - 0.0) / f_sp
        if f >= 1000.0:
            mels = (1000.0 - 0.0) / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
        return mels

    all_freqs = torch.linspace(0, sr // 2, K)
    m_pts = torch.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2)
    if mel_scale == "htk":
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    else:
        f_sp = 200.0 / 3
        f_pts = 0.0 + f_sp * m_pts
        min_log_mel = (1000.0 - 0.0) / f_sp
        logstep = math.log(6.4) / 27.0
        log_t = m_pts >= min_log_mel
        f_pts[log_t] = 1000.0 * torch.exp(logstep * (m_pts[log_t] - min_log_mel))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    if mel_norm:
        fb *= (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb.T.contiguous()


def speech2linear(speech, speech_len, *, n_fft, hop, win, window="hann", center=True, normalized=False,
                  preemphasis=None, mag_spec=False, return_energy=False, logging=False, clamp=1e-10,
                  log_base=None):
    """speech (B,T,1) fp32, speech_len (B,) int64 -> (spec, feat_len, energy|None)."""
    if preemphasis is not None:
        w = torch.tensor([-preemphasis, 1.0]).reshape(1, 1, 2)
        speech = F.conv1d(F.pad(speech.transpose(1, 2), (1, 0)), w).transpose(1, 2)
        for i in range(speech_len.size(0)):
            if speech_len[i] < speech.size(1):
                speech[i][speech_len[i]:] = 0.0
    win_t = getattr(torch, f"{window}_window")(win, dtype=speech.dtype)
    st = torch.stft(speech.squeeze(-1), n_fft=n_fft, hop_length=hop, win_length=win, window=win_t,
                    center=center, normalized=normalized, onesided=True, return_complex=True).transpose(1, 2)
    slen = speech_len + 2 * (n_fft // 2) if center else speech_len
    feat_len = torch.div(slen - n_fft, hop, rounding_mode="floor") + 1
    spec = st.real ** 2 + st.imag ** 2
    energy = torch.sqrt(torch.clamp(spec.sum(dim=-1), min=1e-10)) if return_energy else None
    for i in range(feat_len.size(0)):
        if feat_len[i] < spec.size(1):
            spec[i][feat_len[i]:] = 0.0
            if energy is not None:
                energy[i][feat_len[i]:] = 0.0
    if mag_spec:
        spec = torch.sqrt(spec)
    if logging:
        spec = torch.clamp(spec, min=clamp).log()
        if log_base is not None:
            spec /= math.log(log_base)
    return spec, feat_len, energy


def linear2mel(spec, weight, *, clamp=1e-10, logging=True, log_base=10.0):
    feat = F.linear(spec, weight)
    if logging:
        feat = torch.clamp(feat, min=clamp).log()
        if log_base is not None:
            feat /= math.log(log_base)
    return feat


def utterance_stats(feat, feat_len, clamp=1e-10):
    B = feat.size(0)
    means = torch.stack([feat[i][: feat_len[i]].mean(dim=0) for i in range(B)])
    stds = torch.clamp(torch.stack([feat[i][: feat_len[i]].std(dim=0) for i in range(B)]), min=clamp)
    return means, stds


def norm_utterance(feat, feat_len, clamp=1e-10):
    means, stds = utterance_stats(feat, feat_len, clamp)
    feat = feat - means.unsqueeze(1)
    return feat / stds.unsqueeze(1)


def norm_global_train(feat, feat_len, state: dict, epoch: Optional[int], max_epoch_num=4, clamp=1e-10):
    """Single-process ``norm_type='global'`` training step; ``state`` holds the
    global_batch/global_mean/global_std buffers (feat_norm.py:270-336,455-508)."""
    means, stds = utterance_stats(feat, feat_len, clamp)
    bm, bs = means.mean(dim=0), stds.mean(dim=0)
    if epoch is None or "batch" not in state:
        state["batch"] = torch.LongTensor([1])
    elif epoch <= max_epoch_num:
        state["batch"] = state["batch"] + 1
    if epoch is None or epoch <= max_epoch_num:
        w = 1 / state["batch"]
        state["mean"] = bm if "mean" not in state else w * bm + (1 - w) * state["mean"]
        state["std"] = bs if "std" not in state else w * bs + (1 - w) * state["std"]
    feat = feat - state["mean"]
    return feat / state["std"]


def specaug(feat, feat_len, *, freq_mask_width=(0, 30), freq_mask_num=2, time_mask_width=(0, 0.05),
            time_mask_num=2, feat_norm=True):
    """Masking branch of specaug.py:129-206 (time_warp off) on the global CPU generator."""
    B, D = feat.size(0), feat.size(-1)
    tmax, tmin = feat_len.max().item(), feat_len.min().item()
    ml = torch.randint(freq_mask_width[0], freq_mask_width[1] + 1, size=(B, freq_mask_num)).unsqueeze(2)
    mp = torch.randint(0, max(1, D - ml.max().item()), size=(B, freq_mask_num)).unsqueeze(2)
    ax = torch.arange(D)[None, None, :]
    mask = ((mp <= ax) * (ax < (mp + ml))).any(dim=1, keepdim=True)
    lo, hi = time_mask_width
    lo = int(lo * tmin) if isinstance(lo, float) else lo
    hi = int(hi * tmin) if isinstance(hi, float) else hi
    hi = min(hi, tmin)
    ml = torch.randint(lo, hi + 1, size=(B, time_mask_num)).unsqueeze(1)
    mp = torch.randint(0, max(1, tmin - ml.max().item()), size=(B, time_mask_num)).unsqueeze(1)
    ax = torch.arange(tmax)[None, :, None]
    tm = ((mp <= ax) * (ax < (mp + ml))).any(dim=-1, keepdim=True)
    mask = torch.logical_or(mask, tm)
    return feat.masked_fill(mask, 0.0 if feat_norm else feat.mean())

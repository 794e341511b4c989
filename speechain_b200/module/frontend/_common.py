"""Host-side helpers shared by the frontend modules."""
import math
from typing import Optional

import torch


class ConstHolder(torch.nn.Module):
    """A parameter-free child that only carries a ``weight`` buffer, so the state-dict keeps the
    reference's keys (``…preemph_filter.weight``, ``…mel_fbanks.weight``; speech2linear.py:105-116,
    linear2mel.py:146-155)."""

    def __init__(self, weight: torch.Tensor):
        super().__init__()
        self.register_buffer("weight", weight)

    def extra_repr(self) -> str:
        return f"weight={tuple(self.weight.shape)} (constant buffer)"


def to_samples(x, sr: int) -> int:
    """float => seconds, int => samples (speech2linear.py:85-90)."""
    return int(x * sr) if isinstance(x, float) else x


def as_float(x, name: str) -> Optional[float]:
    """PyYAML parses ``clamp: 1e-5`` as a string (ruamel, which the reference uses, as a float)."""
    if x is None or isinstance(x, float):
        return x
    if isinstance(x, (int, str)):
        try:
            return float(x)
        except ValueError:
            pass
    raise ValueError(f"{name} must be a number, got {x!r}")


def make_window(name, win_length: int, n_fft: int, normalized: bool) -> torch.Tensor:
    """The window torch.stft would use (speech2linear.py:192-196), fp32 on the CPU, with the
    ``normalized=True`` factor n_fft**-0.5 folded in."""
    if isinstance(name, torch.Tensor):
        w = name.detach().to("cpu", torch.float32)
    else:
        fn = getattr(torch, f"{name}_window", None)
        if fn is None:
            raise ValueError(f"unknown window {name!r}: torch.{name}_window does not exist")
        w = fn(win_length, dtype=torch.float32)
    if normalized:
        w = w * float(n_fft) ** -0.5
    return w


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float, mel_scale: str,
                   mel_norm: bool) -> torch.Tensor:
    """(n_mels, n_fft//2+1) fp32 triangular filterbank.

    Restates ``torchaudio.functional.melscale_fbanks(...).T`` (called at linear2mel.py:135-143)
    with the same fp32 op order, so the weights are bit-identical to the reference's buffer
    (checked in tests/test_host_logic.py) without importing torchaudio at run time."""
    n_freqs = n_fft // 2 + 1

    def hz_to_mel(f: float) -> float:
        if mel_scale == "htk":
            return 2595.0 * math.log10(1.0 + (f / 700.0))
        mel = (f - 0.0) / (200.0 / 3)
        if f >= 1000.0:
            mel = (1000.0 - 0.0) / (200.0 / 3) + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
        return mel

    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_pts = torch.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
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
        fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb.T.contiguous()


def squeeze_speech(speech: torch.Tensor) -> torch.Tensor:
    """(B,T,1) or (B,T) -> (B,T) view; 3-D input with a last dim != 1 raises like
    speech2linear.py:158-163."""
    if speech.dim() == 3:
        if speech.size(-1) != 1:
            raise RuntimeError(
                f"Currently, we don't support multi-channel speech waveforms. "
                f"If the speech is given in 3D vectors, the last dimension must be 1. "
                f"But got speech.shape={speech.shape}.")
        speech = speech[..., 0]
    elif speech.dim() != 2:
        raise RuntimeError(f"speech must be (batch, time, 1) or (batch, time), got shape {tuple(speech.shape)}")
    if speech.dtype != torch.float32:
        speech = speech.float()
    return speech


def lengths_pair(length: torch.Tensor, device):
    """(device int64 tensor, host numpy int64 array) for a length tensor living anywhere."""
    from ...host.ops import HostLengths
    if length.dtype != torch.int64:
        length = length.long()
    if length.is_cuda:
        return length, HostLengths.get(length)
    host = length.numpy()
    dev = length.to(device, non_blocking=True)
    HostLengths.remember(dev, host)
    return dev, host

"""Torch-tensor facade over the C ABI (speechain_b200/_lib.py).

PyTorch is plumbing here: it owns device memory and streams, this file turns tensors into raw
pointers and calls libsfe on the current CUDA stream.  No arithmetic on the hot path happens in
torch ops, and nothing here falls back to torch when the library is missing.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from .. import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    # raw handle of torch's current stream on the current device (torch.cuda.current_stream() costs ~8 us)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


class _on_device:
    """`with torch.cuda.device(d)` only when d is not already current (the common case costs nothing)."""

    __slots__ = ("idx", "prev")

    def __init__(self, device: torch.device):
        self.idx = device.index if device.index is not None else torch._C._cuda_getDevice()

    def __enter__(self):
        self.prev = torch._C._cuda_getDevice()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


def _host_len_ptr(len_host, B: int):
    """Address of a host copy of the lengths for the C ABI's optional `wav_len_host` (None when there is none or it
    is not a contiguous int64 array of B entries)."""
    if (isinstance(len_host, np.ndarray) and len_host.dtype == np.int64 and len_host.ndim == 1
            and len_host.shape[0] == B and len_host.flags.c_contiguous):
        return len_host.ctypes.data
    return None


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must live on a CUDA device: the speechain_b200 frontend has no CPU path "
            f"(got device={t.device}).")


class HostLengths:
    """Remembers the host copy (numpy int64) of a device length tensor so that the next module in
    the chain (normalise, SpecAugment) does not pay a device->host sync to learn min/max lengths.

    The host copy rides ON the tensor object (attribute ``_sfe_host_len``, like ``_sfe_stats``) together with
    the tensor's version counter and data pointer: it dies with the tensor, so a fresh length tensor that
    happens to reuse a freed tensor's ``id()`` and allocator block (what the data loader produces every step,
    model/abs.py:329-352) can never see the previous batch's lengths.  A tensor that was modified in place,
    re-pointed (``set_``) or is simply unknown falls back to one ``.cpu()`` copy."""

    _ATTR = "_sfe_host_len"

    @classmethod
    def remember(cls, dev: torch.Tensor, host) -> None:
        """`host`: the numpy int64 copy, or a zero-argument callable that makes it when somebody asks (the chain's
        output lengths: nothing on the hot path reads them on the host)."""
        setattr(dev, cls._ATTR, (dev._version, dev.data_ptr(), host))

    @classmethod
    def get(cls, t: torch.Tensor) -> np.ndarray:
        if not t.is_cuda:
            return t.numpy() if t.dtype == torch.int64 else t.long().numpy()
        hit = getattr(t, cls._ATTR, None)
        if hit is not None and hit[0] == t._version and hit[1] == t.data_ptr():
            host = hit[2]
            if callable(host):
                host = host()
                setattr(t, cls._ATTR, (hit[0], hit[1], host))
            return host
        host = t.cpu().long().numpy()
        cls.remember(t, host)
        return host


class DeviceMasks:
    """SpecAugment interval descriptors.  Small batches travel inside the kernel parameter block
    (host arrays, no copy to the device at all); larger ones are uploaded once per call."""

    def __init__(self, device, freq_pos=None, freq_len=None, time_pos=None, time_len=None, fill: float = 0.0,
                 fill_dev: Optional[torch.Tensor] = None):
        self.n_freq = self.n_time = 0
        self._host = []
        if freq_pos is not None:
            self.n_freq = int(freq_pos.shape[1])
            self._host += [np.ascontiguousarray(freq_pos, dtype=np.int32), np.ascontiguousarray(freq_len, dtype=np.int32)]
        if time_pos is not None:
            self.n_time = int(time_pos.shape[1])
            self._host += [np.ascontiguousarray(time_pos, dtype=np.int32), np.ascontiguousarray(time_len, dtype=np.int32)]
        self.fill = float(fill)
        self.fill_dev = fill_dev
        total = sum(a.size for a in self._host)
        self.on_host = total <= _lib.MASK_INLINE_MAX
        self.buf = None
        if not self.on_host and self._host:
            # (a pageable copy of a few KB: pinning a fresh buffer every step costs five times the copy)
            self.buf = torch.from_numpy(np.concatenate([a.ravel() for a in self._host])).to(device)

    @classmethod
    def persistent(cls, batch_size: int, n_freq: int, n_time: int):
        """Host-resident descriptors whose int32 arrays (and ctypes struct) live as long as the object: the
        SpecAugment draws are written straight into them every step (``views``: torch tensors sharing their
        memory), so a step builds no arrays and no struct.  Only for batches small enough to travel inside the
        kernel parameter block — the library copies them at launch time, before the next draw can touch them."""
        self = cls.__new__(cls)
        self.n_freq, self.n_time = int(n_freq), int(n_time)
        self.fill, self.fill_dev, self.on_host, self.buf = 0.0, None, True, None
        self._host, self.views = [], []
        for n in (n_freq, n_freq, n_time, n_time):
            if n:
                a = np.zeros((batch_size, n), dtype=np.int32)
                self._host.append(a)
                self.views.append(torch.from_numpy(a))
            else:
                self.views.append(None)
        assert sum(a.size for a in self._host) <= _lib.MASK_INLINE_MAX
        # host addresses of (freq_pos, freq_len, time_pos, time_len), 0 where that kind of mask is off
        it = iter(self._host)
        self._addr = tuple(next(it).ctypes.data if n else 0 for n in (n_freq, n_freq, n_time, n_time))
        self._struct = None
        self._struct = self.struct()
        return self

    def struct(self) -> _lib.SfeMasks:
        m = getattr(self, "_struct", None)
        if m is not None:  # persistent descriptors: the pointers never change
            m.fill = self.fill
            m.fill_dev = self.fill_dev.data_ptr() if self.fill_dev is not None else None
            return m
        m = _lib.SfeMasks()
        if self.on_host:
            ptrs = [a.ctypes.data for a in self._host]
        else:
            base, ptrs = self.buf.data_ptr(), []
            for a in self._host:
                ptrs.append(base)
                base += 4 * a.size
        i = 0
        if self.n_freq:
            m.freq_pos, m.freq_len = ptrs[0], ptrs[1]
            i = 2
        if self.n_time:
            m.time_pos, m.time_len = ptrs[i], ptrs[i + 1]
        m.n_freq, m.n_time = self.n_freq, self.n_time
        m.fill = self.fill
        m.fill_dev = self.fill_dev.data_ptr() if self.fill_dev is not None else None
        m.on_host = int(self.on_host)
        return m


class FrontendPlan:
    """Owns one ``sfe_plan`` (constant tables on one device)."""

    def __init__(self, *, n_fft: int, win_length: int, hop_length: int, center: bool = True,
                 preemphasis: Optional[float] = None, mag_spec: bool = False, lin_logging: bool = False,
                 lin_clamp: float = 1e-10, lin_log_base: Optional[float] = None, n_mels: int = 0,
                 mel_logging: bool = True, mel_clamp: float = 1e-10, mel_log_base: Optional[float] = 10.0,
                 window: torch.Tensor = None, mel_weight: Optional[torch.Tensor] = None, device=None):
        lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        cfg = _lib.SfeConfig()
        cfg.n_fft, cfg.win_length, cfg.hop_length, cfg.center = int(n_fft), int(win_length), int(hop_length), int(center)
        cfg.has_preemphasis = int(preemphasis is not None)
        cfg.preemphasis = float(preemphasis or 0.0)
        cfg.mag_spec = int(mag_spec)
        cfg.lin_logging, cfg.lin_clamp = int(lin_logging), float(lin_clamp)
        cfg.lin_log_div = float(math.log(lin_log_base)) if lin_log_base is not None else 1.0
        cfg.n_mels, cfg.mel_logging, cfg.mel_clamp = int(n_mels), int(mel_logging), float(mel_clamp)
        cfg.mel_log_div = float(math.log(mel_log_base)) if mel_log_base is not None else 1.0
        win = window.detach().to("cpu", torch.float32).contiguous()
        if win.numel() != win_length:
            raise ValueError(f"window has {win.numel()} taps, expected win_length={win_length}")
        cfg.window_host = win.data_ptr()
        mw = None
        if n_mels:
            mw = mel_weight.detach().to("cpu", torch.float32).contiguous()
            if tuple(mw.shape) != (n_mels, n_fft // 2 + 1):
                raise ValueError(f"mel_weight has shape {tuple(mw.shape)}, expected {(n_mels, n_fft // 2 + 1)}")
            cfg.mel_weight_host = mw.data_ptr()
        self.n_fft, self.win_length, self.hop_length, self.center = int(n_fft), int(win_length), int(hop_length), bool(center)
        self.n_mels, self.n_bins = int(n_mels), n_fft // 2 + 1
        self.mel_logging = bool(mel_logging)
        self.pad_mel = (np.log(np.float32(mel_clamp)) / np.float32(cfg.mel_log_div)).astype(np.float32).item() \
            if mel_logging else 0.0
        handle = C.c_void_p()
        with _on_device(self.device):
            _lib.check(lib.sfe_plan_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self._lib = lib
        self._ws = {}   # per-batch-size scratch for the fixed-point sums (stream-ordered reuse)
        self._frames = {}  # padded length -> frame count

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.sfe_plan_destroy(h)
            except Exception:
                pass

    def num_frames(self, T_max: int) -> int:
        return int(self._lib.sfe_plan_num_frames(self._h, int(T_max)))

    def query(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.sfe_plan_query(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(smem_bytes=a.value, ctas_per_sm=b.value, n_sm=c.value)

    def forward(self, wav: torch.Tensor, wav_len: torch.Tensor, *, want_mel=True, want_lin=False, want_energy=False,
                want_stats=False, norm_mean=None, norm_std=None, masks: Optional[DeviceMasks] = None,
                skip_pad=False, out_mel=None, out_stats=None, len_host: Optional[np.ndarray] = None):
        """wav (B, T) fp32 CUDA (any row pitch), wav_len (B,) int64 CUDA; len_host: its host copy (contiguous
        numpy int64) when the caller has one — the library then cuts the batch into per-SM ranges on the host."""
        _need_cuda(wav, "speech")
        B, T = wav.shape
        if wav.dtype != torch.float32:
            raise TypeError(f"speech must be float32, got {wav.dtype}")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        F = self.num_frames(T)
        dev = wav.device
        out = {}
        mel = lin = energy = stats = None
        if want_mel:
            mel = out_mel if out_mel is not None else torch.empty((B, F, self.n_mels), dtype=torch.float32, device=dev)
        if want_lin:
            lin = torch.empty((B, F, self.n_bins), dtype=torch.float32, device=dev)
        if want_energy:
            energy = torch.empty((B, F), dtype=torch.float32, device=dev)
        fused_stats = want_stats and self.mel_logging
        if fused_stats:
            # fixed-point sums (see include/sfe.h): int64, recognised downstream by their dtype
            stats = out_stats if out_stats is not None else torch.empty((B, self.n_mels, 2), dtype=torch.int64, device=dev)
        feat_len = torch.empty((B,), dtype=torch.int64, device=dev)
        ms = masks.struct() if masks is not None else None
        with _on_device(dev):
            _lib.check(self._lib.sfe_frontend_forward(
                self._h, _ptr(wav), int(wav.stride(0)), _ptr(wav_len), B, T, _ptr(mel), _ptr(lin), _ptr(energy),
                _ptr(feat_len), _ptr(stats), _ptr(norm_mean), _ptr(norm_std),
                C.byref(ms) if ms is not None else None, int(skip_pad), _stream(), _host_len_ptr(len_host, B)))
        if want_stats and not fused_stats:
            stats = stats_accumulate(mel, feat_len)  # linear-domain mel: general fp64 sums in a second pass
        out.update(mel=mel, lin=lin, energy=energy, stats=stats, feat_len=feat_len)
        return out

    def chain(self, wav: torch.Tensor, wav_len: torch.Tensor, *, norm_mode: int, mean=None, std=None,
              clamp: float = 1e-10, use_mean=True, use_std=True, masks: Optional[DeviceMasks] = None,
              len_host=None, warp=None):
        """frontend -> normalise -> [time warp ->] masks in one library call (see sfe_chain_forward).  Lean on purpose:
        the caller has validated `wav` (B, T) or (B, T, 1) fp32 CUDA with unit stride along T and `wav_len` int64
        CUDA; `len_host` is the host copy of `wav_len` (contiguous numpy int64, or the address of B int64), or None."""
        B, T = wav.shape[0], wav.shape[1]  # (B, T) or (B, T, 1)
        dev = wav.device
        F = self._frames.get(T)
        if F is None:
            F = self._frames[T] = self.num_frames(T)
        mel = torch.empty((B, F, self.n_mels), dtype=torch.float32, device=dev)
        feat_len = torch.empty((B,), dtype=torch.int64, device=dev)
        ws, half = None, 0
        if norm_mode == 2:
            # double-buffered fixed-point sums (include/sfe.h): zeroed once, then the kernels keep them clean
            ent = self._ws.get(B)
            if ent is None or ent[1] < 0:
                ent = self._ws[B] = [torch.zeros((2, B, self.n_mels, 2), dtype=torch.int64, device=dev), 0]
            ws, half = ent
            ent[1] = -1  # marks the scratch dirty until the call below has been enqueued completely
        ms = masks.struct() if masks is not None else None
        wp = out2 = None
        if warp is not None:  # the last pass runs out of place: a second buffer takes the result
            wp = _lib.SfeWarp(int(warp[3]) if len(warp) > 3 else 1, int(warp[0]), int(warp[1]), int(warp[2]))
            out2 = torch.empty((B, F, self.n_mels), dtype=torch.float32, device=dev)
        idx = dev.index
        prev = torch._C._cuda_getDevice()
        if idx is not None and idx != prev:
            torch.cuda.set_device(idx)
        else:
            idx = prev
        try:
            rc = self._lib.sfe_chain_forward(
                self._h, wav.data_ptr(), wav.stride(0), wav_len.data_ptr(), B, T, mel.data_ptr(), feat_len.data_ptr(),
                ws.data_ptr() if ws is not None else None, half, norm_mode,
                mean.data_ptr() if mean is not None else None, std.data_ptr() if std is not None else None,
                clamp, int(use_mean), int(use_std), None, None, C.byref(ms) if ms is not None else None,
                torch._C._cuda_getCurrentRawStream(idx),
                len_host if isinstance(len_host, int) else _host_len_ptr(len_host, B),
                C.byref(wp) if wp is not None else None, out2.data_ptr() if out2 is not None else None)
        finally:
            if idx != prev:
                torch.cuda.set_device(prev)
        if rc:
            _lib.check(rc)
        if ws is not None:
            self._ws[B][1] = 1 - half
        return (out2 if out2 is not None else mel), feat_len

    def chain_sums(self, wav: torch.Tensor, wav_len: torch.Tensor, *, clamp: float = 1e-10, len_host=None):
        """First half of a step whose batch / global statistics are still being averaged (sfe_chain_forward, norm_mode
        3): raw valid rows + the packed [sum of means | sum of stds | B] of the batch.  Same lean contract as `chain`.
        Returns (mel, feat_len, packed (2 n_mels + 1,) fp32)."""
        B, T = wav.shape[0], wav.shape[1]
        dev = wav.device
        F = self._frames.get(T)
        if F is None:
            F = self._frames[T] = self.num_frames(T)
        mel = torch.empty((B, F, self.n_mels), dtype=torch.float32, device=dev)
        feat_len = torch.empty((B,), dtype=torch.int64, device=dev)
        packed = torch.empty((2 * self.n_mels + 1,), dtype=torch.float32, device=dev)
        # double-buffered fixed-point sums, shared with norm_mode 2 (same protocol: see `chain`)
        ent = self._ws.get(B)
        if ent is None or ent[1] < 0:
            ent = self._ws[B] = [torch.zeros((2, B, self.n_mels, 2), dtype=torch.int64, device=dev), 0]
        ws, half = ent
        ent[1] = -1  # marks the scratch dirty until the call below has been enqueued completely
        idx = dev.index
        prev = torch._C._cuda_getDevice()
        if idx is not None and idx != prev:
            torch.cuda.set_device(idx)
        else:
            idx = prev
        try:
            rc = self._lib.sfe_chain_forward(
                self._h, wav.data_ptr(), wav.stride(0), wav_len.data_ptr(), B, T, mel.data_ptr(), feat_len.data_ptr(),
                ws.data_ptr(), half, 3, None, None, clamp, 1, 1, packed.data_ptr(), None, None,
                torch._C._cuda_getCurrentRawStream(idx),
                len_host if isinstance(len_host, int) else _host_len_ptr(len_host, B), None, None)
        finally:
            if idx != prev:
                torch.cuda.set_device(prev)
        if rc:
            _lib.check(rc)
        self._ws[B][1] = 1 - half
        return mel, feat_len, packed

    def griffinlim(self, spec: torch.Tensor, feat_len: torch.Tensor, *, n_iter: int = 32, momentum: float = 0.99,
                   init_angles: Optional[torch.Tensor] = None, inv_preemph_taps: int = 0):
        """(B, F, K) power/magnitude spectrogram -> (waveform (B, hop*(F-1)), wav_len (B,)); see sfe_griffinlim.
        `init_angles` (B, F, K) complex64 on the device, or None for an all-ones start."""
        _need_cuda(spec, "feat")
        spec = spec.contiguous()
        B, F, K = spec.shape
        if K != self.n_bins:
            raise RuntimeError(f"expected {self.n_bins} linear bins, got feat.shape={tuple(spec.shape)}")
        dev = spec.device
        ang = None
        if init_angles is not None:
            ang = torch.view_as_real(init_angles.to(dev, torch.complex64).contiguous())
        nbytes = self._lib.sfe_griffinlim_workspace(self._h, B, F)
        if nbytes < 0:
            raise RuntimeError("sfe_griffinlim_workspace failed")
        work = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=dev)
        wav = torch.empty((B, self.hop_length * max(F - 1, 0)), dtype=torch.float32, device=dev)
        wav_len = torch.empty((B,), dtype=torch.int64, device=dev)
        with _on_device(dev):
            _lib.check(self._lib.sfe_griffinlim(self._h, _ptr(spec), _ptr(feat_len), B, F, int(n_iter), float(momentum),
                                                _ptr(ang), int(inv_preemph_taps), _ptr(work), int(nbytes), _ptr(wav),
                                                _ptr(wav_len), _stream()))
        return wav, wav_len

    def linear2mel(self, lin: torch.Tensor) -> torch.Tensor:
        _need_cuda(lin, "feat")
        lin = lin.contiguous()
        rows = lin.numel() // self.n_bins
        mel = torch.empty(lin.shape[:-1] + (self.n_mels,), dtype=torch.float32, device=lin.device)
        with _on_device(lin.device):
            _lib.check(self._lib.sfe_linear2mel_forward(self._h, _ptr(lin), rows, _ptr(mel), _stream()))
        return mel


# ----------------------------------------------------------------------------------------------
# plan-free entry points
# ----------------------------------------------------------------------------------------------

def stats_accumulate(feat: torch.Tensor, feat_len: torch.Tensor) -> torch.Tensor:
    _need_cuda(feat, "feat")
    B, F, D = feat.shape
    stats = torch.empty((B, D, 2), dtype=torch.float64, device=feat.device)
    with _on_device(feat.device):
        _lib.check(_lib.load().sfe_stats_accumulate(_ptr(feat), _ptr(feat_len), B, F, D, _ptr(stats), _stream()))
    return stats


def stats_finalize(stats: torch.Tensor, feat_len: torch.Tensor, clamp: float, want_packed: bool):
    B, D, _ = stats.shape
    dev = stats.device
    mean = torch.empty((B, D), dtype=torch.float32, device=dev)
    std = torch.empty((B, D), dtype=torch.float32, device=dev)
    packed = torch.empty((2 * D + 1,), dtype=torch.float32, device=dev) if want_packed else None
    with _on_device(dev):
        _lib.check(_lib.load().sfe_stats_finalize(_ptr(stats), int(stats.dtype == torch.int64), _ptr(feat_len), B, D,
                                                  float(clamp), _ptr(mean), _ptr(std), _ptr(packed), _stream()))
    return mean, std, packed


def running_update(packed: torch.Tensor, D: int, weight: float, first: bool, update: bool, run_mean, run_std,
                   want_batch: bool = False):
    dev = packed.device
    bm = torch.empty((D,), dtype=torch.float32, device=dev) if want_batch else None
    bs = torch.empty((D,), dtype=torch.float32, device=dev) if want_batch else None
    with _on_device(dev):
        _lib.check(_lib.load().sfe_running_update(_ptr(packed), D, float(weight), int(first), int(update), _ptr(run_mean),
                                                  _ptr(run_std), _ptr(bm), _ptr(bs), _stream()))
    return bm, bs


def running_normalize(packed: torch.Tensor, weight: float, first: bool, run_mean, run_std, counter, count: int,
                      feat: torch.Tensor, feat_len: torch.Tensor, *, src_has_pad: bool, pad_value: float,
                      masks: Optional["DeviceMasks"] = None, warp=None, dst: Optional[torch.Tensor] = None):
    """Second half of the averaging step (sfe_running_normalize): running average of the statistics from the packed
    sums, `count` into the counter buffer, then (x - mean) / std [+ time warp] [+ masks] over `feat` into `dst` — in
    place when `dst` is None and no warp is requested (a warp always gets a new tensor).  Returns the result tensor."""
    B, F, D = feat.shape
    dev = feat.device
    wp = None
    if warp is not None:
        wp = _lib.SfeWarp(int(warp[3]) if len(warp) > 3 else 1, int(warp[0]), int(warp[1]), int(warp[2]))
        if dst is None or dst is feat:
            dst = torch.empty_like(feat)
    elif dst is None:
        dst = feat
    ms = masks.struct() if masks is not None else None
    with _on_device(dev):
        _lib.check(_lib.load().sfe_running_normalize(
            packed.data_ptr(), D, float(weight), int(first), run_mean.data_ptr() if run_mean is not None else None,
            run_std.data_ptr() if run_std is not None else None, counter.data_ptr() if counter is not None else None,
            int(count), feat.data_ptr(), feat_len.data_ptr(), B, F, int(src_has_pad), float(pad_value),
            C.byref(ms) if ms is not None else None, _stream(), C.byref(wp) if wp is not None else None,
            dst.data_ptr()))
    return dst


def normalize(src: torch.Tensor, feat_len: torch.Tensor, *, dst: Optional[torch.Tensor] = None, stats=None,
              clamp: float = 1e-10, use_mean=True, use_std=True, mean=None, std=None, per_utt=False,
              want_utt_stats=False, src_has_pad=True, pad_value=0.0, warp=None, masks: Optional[DeviceMasks] = None,
              out_bf16: bool = False):
    """(x-mean)/std [+ warp] [+ masks]; in place unless ``dst`` is given or a warp is requested.
    ``out_bf16``: the result is written as a new bfloat16 tensor (sfe_normalize_bf16) — the dtype an AMP prenet takes."""
    _need_cuda(src, "feat")
    B, F, D = src.shape
    if out_bf16:
        if want_utt_stats:
            raise ValueError("per-utterance statistics outputs are not available with out_bf16")
        out = torch.empty((B, F, D), dtype=torch.bfloat16, device=src.device)
        w = _lib.SfeWarp(int(warp[3]) if len(warp) > 3 else 1, int(warp[0]), int(warp[1]), int(warp[2])) if warp is not None else None
        ms = masks.struct() if masks is not None else None
        with _on_device(src.device):
            _lib.check(_lib.load().sfe_normalize_bf16(
                _ptr(src), _ptr(out), _ptr(feat_len), B, F, D, _ptr(stats),
                int(stats is not None and stats.dtype == torch.int64), float(clamp), int(use_mean), int(use_std),
                _ptr(mean), _ptr(std), int(per_utt), int(src_has_pad), float(pad_value),
                C.byref(w) if w is not None else None, C.byref(ms) if ms is not None else None, _stream()))
        return out, None, None
    if warp is not None and dst is None:
        dst = torch.empty_like(src)
    if dst is None:
        dst = src
    mo = so = None
    if want_utt_stats and stats is not None:
        mo = torch.empty((B, D), dtype=torch.float32, device=src.device)
        so = torch.empty((B, D), dtype=torch.float32, device=src.device)
    w = None
    if warp is not None:
        w = _lib.SfeWarp(int(warp[3]) if len(warp) > 3 else 1, int(warp[0]), int(warp[1]), int(warp[2]))
    ms = masks.struct() if masks is not None else None
    with _on_device(src.device):
        _lib.check(_lib.load().sfe_normalize(
            _ptr(src), _ptr(dst), _ptr(feat_len), B, F, D, _ptr(stats),
            int(stats is not None and stats.dtype == torch.int64), float(clamp), int(use_mean), int(use_std),
            _ptr(mean), _ptr(std), int(per_utt), _ptr(mo), _ptr(so), int(src_has_pad), float(pad_value),
            C.byref(w) if w is not None else None, C.byref(ms) if ms is not None else None, _stream()))
    return dst, mo, so


def prestft_norm(wav: torch.Tensor, wav_len: torch.Tensor, *, preemphasis: Optional[float], mode: str) -> torch.Tensor:
    """Pre-emphasis + zero beyond each length + waveform normalisation over the whole padded row
    (speech2linear.py:165-187) -> new (B, T) fp32 tensor; see sfe_prestft_norm."""
    _need_cuda(wav, "speech")
    B, T = wav.shape
    if mode == "mean_std" and B != 1:
        raise RuntimeError("pre_stft_norm='mean_std' only broadcasts for batch size 1 (speech2linear.py:179 has no keepdim)")
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    out = torch.empty((B, T), dtype=torch.float32, device=wav.device)
    scratch = torch.empty((B, 2), dtype=torch.float64, device=wav.device)
    with _on_device(wav.device):
        _lib.check(_lib.load().sfe_prestft_norm(
            _ptr(wav), int(wav.stride(0)), _ptr(wav_len), B, T, int(preemphasis is not None),
            float(preemphasis or 0.0), 1 if mode == "mean_std" else 2, _ptr(scratch), _ptr(out), T, _stream()))
    return out


def twosided(power: torch.Tensor, feat_len: torch.Tensor, n_fft: int, *, mag_spec: bool, logging: bool, clamp: float,
             log_base: Optional[float], want_energy: bool):
    """One-sided power spectrogram (B, F, n_fft//2+1) -> (two-sided spectrogram (B, F, n_fft) after the reference's
    point-wise tail, frame energy (B, F) over all n_fft bins or None); see sfe_twosided."""
    _need_cuda(power, "feat")
    power = power.contiguous()
    B, F, K = power.shape
    out = torch.empty((B, F, n_fft), dtype=torch.float32, device=power.device)
    energy = torch.empty((B, F), dtype=torch.float32, device=power.device) if want_energy else None
    with _on_device(power.device):
        _lib.check(_lib.load().sfe_twosided(_ptr(power), _ptr(feat_len), B, F, K, int(n_fft), int(mag_spec), int(logging),
                                            float(clamp), float(math.log(log_base)) if log_base is not None else 1.0,
                                            _ptr(out), _ptr(energy), _stream()))
    return out, energy


def delta(feat: torch.Tensor, order: int, N: int) -> torch.Tensor:
    _need_cuda(feat, "feat")
    feat = feat.contiguous()
    B, F, D = feat.shape
    out = torch.empty((B, F, D * (order + 1)), dtype=torch.float32, device=feat.device)
    with _on_device(feat.device):
        _lib.check(_lib.load().sfe_delta(_ptr(feat), B, F, D, int(order), int(N), _ptr(out), _stream()))
    return out


def mean_all(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, "feat")
    x = x.contiguous()
    out = torch.empty((1,), dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _lib.check(_lib.load().sfe_mean_all(_ptr(x), x.numel(), _ptr(out), _stream()))
    return out


def upload_waveforms(wav_host: torch.Tensor, len_host: torch.Tensor, out: Optional[torch.Tensor] = None,
                     device=None) -> torch.Tensor:
    """Ragged H2D copy of a collated batch: only the valid prefix of every row crosses PCIe (rows of similar
    length travel as one pitched copy, dragging along at most 1/24 of padding).

    wav_host (B, T, 1) or (B, T) fp32 CPU (pinned for an asynchronous copy), len_host (B,) int64 CPU.
    Returns the device tensor of the same shape; samples beyond each length are NOT initialised
    (the frontend never reads them).  Replaces the padded copy of model/abs.py:329-352 for this path."""
    if wav_host.is_cuda or len_host.is_cuda:
        raise RuntimeError("upload_waveforms expects host tensors")
    if wav_host.dtype != torch.float32 or not wav_host.is_contiguous():
        wav_host = wav_host.float().contiguous()
    len_host = len_host.long().contiguous()
    B, T = wav_host.shape[0], wav_host.shape[1]
    if out is None:
        out = torch.empty(wav_host.shape, dtype=torch.float32, device=device if device is not None else "cuda")
    with _on_device(out.device):
        _lib.check(_lib.load().sfe_upload_waveforms(wav_host.data_ptr(), T, len_host.data_ptr(), B, out.data_ptr(), T, _stream()))
    return out


def upload_waveforms_i16(wav_host: torch.Tensor, len_host: torch.Tensor, out: Optional[torch.Tensor] = None,
                         stage: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """``upload_waveforms`` for 16-bit PCM: wav_host (B, T, 1) or (B, T) int16 CPU (pinned for an asynchronous copy) —
    the samples as they sit in the corpus files — cross PCIe at two bytes each and become fp32 ``x / 32768`` on the
    device, bit for bit what ``soundfile.read(dtype="float32")`` hands the reference's data loader
    (utilbox/data_loading_util.py:71-78).  ``stage`` is an optional reusable uint8 device buffer of at least
    ``upload_i16_stage_bytes(B, T)`` bytes; returns the fp32 device tensor (samples beyond each length NOT initialised)."""
    if wav_host.is_cuda or len_host.is_cuda:
        raise RuntimeError("upload_waveforms_i16 expects host tensors")
    if wav_host.dtype != torch.int16:
        raise TypeError(f"upload_waveforms_i16 expects int16 PCM, got {wav_host.dtype}")
    if not wav_host.is_contiguous():
        wav_host = wav_host.contiguous()
    len_host = len_host.long().contiguous()
    B, T = wav_host.shape[0], wav_host.shape[1]
    if out is None:
        out = torch.empty(wav_host.shape, dtype=torch.float32, device=device if device is not None else "cuda")
    lib = _lib.load()
    need = int(lib.sfe_upload_i16_stage_bytes(B, T))
    if stage is None or stage.numel() < need:
        stage = torch.empty((need,), dtype=torch.uint8, device=out.device)
    with _on_device(out.device):
        _lib.check(lib.sfe_upload_waveforms_i16(wav_host.data_ptr(), T, len_host.data_ptr(), B, stage.data_ptr(),
                                                stage.numel(), out.data_ptr(), T, _stream()))
    return out


def upload_i16_stage_bytes(B: int, T: int) -> int:
    return int(_lib.load().sfe_upload_i16_stage_bytes(B, T))


def upload_bytes(len_host: torch.Tensor) -> int:
    """Bytes ``upload_waveforms`` moves for these lengths (valid samples + the padding inside its row groups)."""
    len_host = len_host.long().contiguous()
    return int(_lib.load().sfe_upload_bytes(len_host.data_ptr(), len_host.numel()))


def mel2linear(mel: torch.Tensor, feat_len: torch.Tensor, pinv_t: torch.Tensor, *, log_mode: int, log_base: float,
               max_len: int, clamp: float = 1e-10) -> torch.Tensor:
    """(B, F, D) (log-)mel -> (B, F, K) linear spectrogram through pinv(W)^T (D, K); see sfe_mel2linear."""
    _need_cuda(mel, "feat")
    mel = mel.contiguous()
    B, F, D = mel.shape
    K = pinv_t.shape[1]
    lin = torch.empty((B, F, K), dtype=torch.float32, device=mel.device)
    with _on_device(mel.device):
        _lib.check(_lib.load().sfe_mel2linear(_ptr(mel), _ptr(feat_len), B, F, D, K, _ptr(pinv_t), int(log_mode),
                                              float(log_base), int(max_len), float(clamp), _ptr(lin), _stream()))
    return lin

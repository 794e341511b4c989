"""Batch-level SpecAugment.  API mirror of speechain/module/augment/specaug.py (`SpecAugment`).

The random draws are made on the HOST with the reference's exact ``torch.randint`` call sequence
on torch's CPU generator (specaug.py:98-108,145-158,178-191), so masks are bit-identical to the
reference CPU path under the same ``torch.manual_seed``; only interval descriptors (a few hundred
bytes) go to the device, where one kernel applies the time warp (4-tap cubic gather) and the
masks in a single read + write."""
from typing import List, Optional, Union

import torch

from ...host import ops
from ...host.ops import DeviceMasks, HostLengths
from ..abs import Module


_DRAW = [False, None]  # [looked for, module]


def _draw_helper():
    """The optional host helper (csrc/sfe_draw.cpp, built by speechain_b200.csrc.build.build_draw_ext): the draws of a
    step in one call on torch's default CPU generator.  None where it has not been built — the tensor-op path below
    makes the same draws.  SFE_DRAW_HELPER=0 switches it off (A/B runs, tests of the tensor-op path)."""
    if not _DRAW[0]:
        _DRAW[0] = True
        import os
        if os.environ.get("SFE_DRAW_HELPER", "1") != "0":
            try:
                import importlib.util
                from ...csrc import build as _b
                path = _b.draw_ext_path()
                if os.path.exists(path):
                    spec = importlib.util.spec_from_file_location(_b.DRAW_NAME, path)
                    mod = importlib.util.module_from_spec(spec)
                    spec.loader.exec_module(mod)
                    _DRAW[1] = mod
            except Exception:  # noqa: BLE001  (host glue only: fall back to the tensor-op draws)
                _DRAW[1] = None
    return _DRAW[1]


class SpecAugment(Module):
    def module_init(self, time_warp: bool = True, time_warp_window: int = 5, time_warp_mode: str = "bicubic",
                    freq_mask: bool = True, freq_mask_width: Union[int, List[int]] = 30, freq_mask_num: int = 2,
                    time_mask: bool = True, time_mask_width: Union[int, float, List] = 0.05, time_mask_num: int = 2,
                    time_mask_ratio: float = 1.0, feat_norm: bool = True):
        assert time_warp or freq_mask or time_mask, "You must specify at least one type of augmentation in SpecAugment!"
        self.feat_dim = None
        if self.input_size is not None:
            self.feat_dim = self.input_size
            self.output_size = self.input_size
        self.time_warp = time_warp
        self.time_warp_window = time_warp_window
        self.time_warp_mode = time_warp_mode
        # F.interpolate(..., align_corners=True) on the (B, 1, T, D) tensor (specaug.py:111-122) only accepts these two
        if time_warp and time_warp_mode not in ("bicubic", "bilinear"):
            raise ValueError(f"time_warp_mode must be 'bicubic' or 'bilinear' (align_corners=True interpolation of a "
                             f"4-D tensor), got {time_warp_mode!r}")
        self.freq_mask = freq_mask
        if isinstance(freq_mask_width, int):
            freq_mask_width = [0, freq_mask_width]
        elif not isinstance(freq_mask_width, (list, tuple)):
            raise ValueError
        if self.feat_dim is not None:
            assert freq_mask_width[1] < self.feat_dim, (
                "The number of maximum frequency masking bins cannot be larger than the feature dimension! "
                f"freq_mask_width[1]={freq_mask_width[1]} and self.feat_dim={self.feat_dim}.")
        self.freq_mask_width = list(freq_mask_width)
        self.freq_mask_num = freq_mask_num
        self.time_mask = time_mask
        if isinstance(time_mask_width, (int, float)):
            time_mask_width = [0, time_mask_width]
        elif not isinstance(time_mask_width, (list, tuple)):
            raise ValueError
        self.time_mask_width = list(time_mask_width)
        self.time_mask_num = time_mask_num
        self.feat_norm = feat_norm

    def __getstate__(self):
        # the persistent descriptor arrays hold raw host addresses: never pickle / deepcopy them
        d = self.__dict__.copy()
        d.pop("_mask_cache", None)
        d.pop("_warp_draw", None)
        return d

    # ------------------------------------------------------------------------------------
    def draw(self, batch_size: int, feat_dim: int, time_minlen: int, generator: Optional[torch.Generator] = None):
        """The reference's draw sequence on the CPU generator.  Returns
        (warp or None, freq_pos, freq_len, time_pos, time_len) with numpy int64 arrays / None."""
        kw = {} if generator is None else {"generator": generator}
        warp = None
        if self.time_warp and time_minlen > 2 * self.time_warp_window + 1:
            w = self.time_warp_window
            center = torch.randint(low=w + 1, high=time_minlen - w, size=(1,), **kw)[0].item()
            pos = torch.randint(low=center - w, high=center + w, size=(1,), **kw)[0].item()
            warp = (center, pos, time_minlen, 2 if self.time_warp_mode == "bilinear" else 1)
        fpos = flen = tpos = tlen = None
        if self.freq_mask:
            if self.feat_dim is None:
                # latent bug of specaug.py:135-141, kept as the same error: feat_dim is never set lazily
                assert self.feat_dim == feat_dim and self.freq_mask_width[1] < feat_dim, (
                    "The number of maximum frequency masking bins cannot be larger than the feature dimension! "
                    f"self.freq_mask_width[1]={self.freq_mask_width[1]} and feat_dim={feat_dim}.")
            flen = torch.randint(self.freq_mask_width[0], self.freq_mask_width[1] + 1,
                                 size=(batch_size, self.freq_mask_num), **kw).numpy()
            # (the maximum through numpy: same value as flen.max().item(), a third of the host time)
            fpos = torch.randint(0, max(1, feat_dim - int(flen.max())), size=(batch_size, self.freq_mask_num), **kw)
            fpos = fpos.numpy()
        if self.time_mask:
            lo, hi = self.time_mask_width
            if isinstance(lo, float):
                lo = int(lo * time_minlen)
            if isinstance(hi, float):
                hi = int(hi * time_minlen)
            hi = min(hi, time_minlen)
            tlen = torch.randint(lo, hi + 1, size=(batch_size, self.time_mask_num), **kw).numpy()
            tpos = torch.randint(0, max(1, time_minlen - int(tlen.max())), size=(batch_size, self.time_mask_num), **kw)
            tpos = tpos.numpy()
        return warp, fpos, flen, tpos, tlen

    def draw_device(self, batch_size, feat_dim, feat_len_host, device, generator=None, minlen=None):
        """Draw and hand to the device: returns (warp, DeviceMasks or None).  ``minlen``: min of the feature lengths
        when the caller already has it (``feat_len_host`` may then be None).

        Small batches (the descriptors fit the kernel parameter block) take a lean path: the draws of ``draw`` —
        same order, bounds and shapes, hence the same stream — are written as int32 straight into persistent host
        arrays, whose ctypes descriptor is built once: ``t.random_(low, high)`` is what ``torch.randint(low, high,
        size, out=t)`` runs after parsing its arguments (ATen's randint_out: resize, then random_(low, high)), two
        thirds of its host time.  The returned object is valid until the next draw of this module."""
        if minlen is None:
            minlen = int(feat_len_host.min())
        nf = self.freq_mask_num if self.freq_mask else 0
        nt = self.time_mask_num if self.time_mask else 0
        if generator is not None or not 0 < 2 * batch_size * (nf + nt) <= ops._lib.MASK_INLINE_MAX:
            warp, fpos, flen, tpos, tlen = self.draw(batch_size, feat_dim, minlen, generator)
            masks = None
            if fpos is not None or tpos is not None:
                masks = DeviceMasks(device, fpos, flen, tpos, tlen, fill=0.0)
            return warp, masks
        cache = self.__dict__.setdefault("_mask_cache", {})
        masks = cache.get(batch_size)
        if masks is None:
            masks = cache[batch_size] = DeviceMasks.persistent(batch_size, nf, nt)
        masks.fill, masks.fill_dev = 0.0, None
        if nf and self.feat_dim is None:
            # latent bug of specaug.py:135-141, kept as the same error: feat_dim is never set lazily
            assert self.feat_dim == feat_dim and self.freq_mask_width[1] < feat_dim, (
                "The number of maximum frequency masking bins cannot be larger than the feature dimension! "
                f"self.freq_mask_width[1]={self.freq_mask_width[1]} and feat_dim={feat_dim}.")
        lo = hi = 0
        if nt:
            lo, hi = self.time_mask_width
            if isinstance(lo, float):
                lo = int(lo * minlen)
            if isinstance(hi, float):
                hi = int(hi * minlen)
            hi = min(hi, minlen)
        helper = _draw_helper()
        if helper is not None:
            # every draw of this step in ONE call on the same generator (csrc/sfe_draw.cpp)
            ptr = masks._addr
            center, pos = helper.specaug_draw(
                batch_size, feat_dim, minlen, self.time_warp_window if self.time_warp else -1,
                nf, self.freq_mask_width[0] if nf else 0, self.freq_mask_width[1] if nf else 0, ptr[0], ptr[1],
                nt, lo, hi, ptr[2], ptr[3])
            warp = (center, pos, minlen, 2 if self.time_warp_mode == "bilinear" else 1) if center >= 0 else None
            return warp, masks
        warp = None
        if self.time_warp and minlen > 2 * self.time_warp_window + 1:
            # (the two scalar draws of specaug.py:98-108 — randint(size=(1,)) — in place on a persistent tensor)
            w = self.time_warp_window
            t1 = self.__dict__.get("_warp_draw")
            if t1 is None:
                t1 = self.__dict__["_warp_draw"] = torch.empty((1,), dtype=torch.int64)
            center = t1.random_(w + 1, minlen - w).item()
            pos = t1.random_(center - w, center + w).item()
            warp = (center, pos, minlen, 2 if self.time_warp_mode == "bilinear" else 1)
        t_fpos, t_flen, t_tpos, t_tlen = masks.views
        if nf:
            t_flen.random_(self.freq_mask_width[0], self.freq_mask_width[1] + 1)
            t_fpos.random_(0, max(1, feat_dim - int(masks._host[1].max())))
        if nt:
            t_tlen.random_(lo, hi + 1)
            t_tpos.random_(0, max(1, minlen - int(masks._host[-1].max())))
        return warp, masks

    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor):
        """feat (batch, F, D) fp32 CUDA, feat_len (batch,).  Returns (augmented feat, feat_len);
        the input tensor is not modified (the reference writes the warp in place, specaug.py:123-124)."""
        if not feat.is_cuda:
            raise RuntimeError("speechain_b200 SpecAugment: `feat` must be a CUDA tensor — there is no CPU path.")
        if feat.dtype != torch.float32 or not feat.is_contiguous():
            feat = feat.float().contiguous()
        B, _, D = feat.shape
        len_host = HostLengths.get(feat_len)
        if not feat_len.is_cuda:
            feat_len = feat_len.to(feat.device, non_blocking=True)
        warp, masks = self.draw_device(B, D, len_host, feat.device)
        if masks is None and warp is None:
            return feat, feat_len
        if masks is not None and not self.feat_norm:
            # fill = mean of the (warped) features (specaug.py:203): warp first, then reduce, then mask
            if warp is not None:
                feat, _, _ = ops.normalize(feat, feat_len, warp=warp)
                warp = None
            masks.fill_dev = ops.mean_all(feat)
        out, _, _ = ops.normalize(feat, feat_len, dst=torch.empty_like(feat), warp=warp, masks=masks)
        return out, feat_len

    def extra_repr(self) -> str:
        out = ""
        if self.time_warp:
            out += f"time_warp_window={self.time_warp_window}, time_warp_mode={self.time_warp_mode}, "
        if self.freq_mask:
            out += f"\nfreq_mask_width={self.freq_mask_width}, freq_mask_num={self.freq_mask_num}, "
        if self.time_mask:
            out += f"\ntime_mask_width={self.time_mask_width}, time_mask_num={self.time_mask_num}"
        return out

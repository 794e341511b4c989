"""Fused frontend -> normalise -> SpecAugment chain.

The reference's ASREncoder calls the three modules one after the other (encoder/asr.py:103-119);
each is a drop-in here too, but calling them separately costs one extra read+write of the feature
tensor per module.  ``AcousticFrontend`` owns the three and issues the cheapest schedule:

* statistics known before the batch (eval, or global statistics frozen after ``max_epoch_num``):
  ONE kernel — normalisation and masks run in the epilogue of the fused frontend kernel and the
  feature tensor is written exactly once;
* statistics that depend on the batch (utterance / batch / global-while-averaging): the frontend
  kernel writes the raw valid frames once and accumulates fp64 sums, then one normalise(+warp)
  (+mask) pass finishes in place — two launches (plus the 2·D+1-float all-reduce for DDP).
"""
from typing import Optional

import torch

from ..host import ops
from .abs import Module
from .augment.specaug import SpecAugment
from .frontend.speech2mel import Speech2MelSpec
from .norm.feat_norm import FeatureNormalization


class AcousticFrontend(Module):
    def module_init(self, frontend: dict, normalize=None, specaug: Optional[dict] = None):
        """``frontend`` / ``normalize`` / ``specaug`` take the same YAML blocks as ASREncoder
        (encoder/asr.py:44-68): frontend = {type:…, conf:{…}} or the conf dict itself;
        normalize = True | dict | None; specaug = dict | None."""
        conf = frontend.get("conf", frontend) if isinstance(frontend, dict) else {}
        self.frontend = Speech2MelSpec(**conf)
        self.frontend.collect_stats = False
        size = self.frontend.output_size
        if normalize is not None and normalize is not False:
            nconf = {} if normalize is True else dict(normalize)
            self.normalize = FeatureNormalization(input_size=size, distributed=self.distributed, **nconf)
        if specaug is not None and specaug is not False:
            sconf = {} if specaug is True else dict(specaug)
            self.specaug = SpecAugment(input_size=size, feat_norm=hasattr(self, "normalize"), **sconf)
        self.output_size = size

    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor, epoch: int = None, domain=None):
        fe = self.frontend
        norm: Optional[FeatureNormalization] = getattr(self, "normalize", None)
        aug: Optional[SpecAugment] = getattr(self, "specaug", None) if self.training else None
        if fe.delta_order is not None or fe.return_energy:
            return self._unfused(feat, feat_len, epoch, domain)
        if norm is not None and not (norm.mean_norm or norm.std_norm):
            norm = None

        # SpecAugment draws need only host-side lengths, so they happen before any launch
        warp = masks = None
        if aug is not None:
            wav_len_host = ops.HostLengths.get(feat_len)
            flen_host = fe.speech2linear._feat_len_host(wav_len_host.long())
            warp, masks = aug.draw_device(feat.size(0), fe.output_size, flen_host, feat.device)
            if masks is not None and not aug.feat_norm:
                return self._unfused(feat, feat_len, epoch, domain, drawn=(warp, masks))

        if norm is None:
            out = fe.run(feat, feat_len, masks=masks if warp is None else None)
            y, ylen = out["mel"], out["feat_len"]
            if warp is not None:
                y, _, _ = ops.normalize(y, ylen, warp=warp, masks=masks)
            return y, ylen

        kind = norm.norm_type
        if kind == "group":
            assert domain is not None, "group-level normalisation needs group_ids (`domain`)"
        prefix = "global" if kind == "global" else (str(domain) if isinstance(domain, (str, int)) else None)
        known = (kind in ("global", "group") and prefix is not None and not isinstance(domain, torch.Tensor)
                 and (not norm.training or (epoch is not None and epoch > norm.max_epoch_num))
                 and (not norm.mean_norm or hasattr(norm, f"{prefix}_mean"))
                 and (not norm.std_norm or hasattr(norm, f"{prefix}_std")))
        if known:
            # single pass: (x-mean)/std and masks in the frontend kernel's epilogue
            if norm.training:
                norm._counter(prefix)  # (frozen counter, nothing to update; feat_norm.py:477-483)
            mean = norm.get_buffer(f"{prefix}_mean") if norm.mean_norm else None
            std = norm.get_buffer(f"{prefix}_std") if norm.std_norm else None
            out = fe.run(feat, feat_len, norm_mean=mean, norm_std=std, masks=masks if warp is None else None)
            y, ylen = out["mel"], out["feat_len"]
            if warp is not None:
                y, _, _ = ops.normalize(y, ylen, warp=warp, masks=masks)
            return y, ylen

        # two passes: raw valid frames + sums, then normalise(+warp)(+mask)
        out = fe.run(feat, feat_len, want_stats=True, skip_pad=True)
        y, ylen = out["mel"], out["feat_len"]
        y._sfe_stats = (y._version, out["stats"])
        if warp is None:
            y, ylen = norm(y, ylen, group_ids=domain, epoch=epoch, masks=masks, inplace=True, src_has_pad=False,
                           pad_value=out["pad_value"])
        else:
            # the warp gathers neighbouring rows, so it cannot share the in-place normalise pass
            y, ylen = norm(y, ylen, group_ids=domain, epoch=epoch, inplace=True, src_has_pad=False,
                           pad_value=out["pad_value"])
            y, _, _ = ops.normalize(y, ylen, warp=warp, masks=masks)
        return y, ylen

    def _unfused(self, feat, feat_len, epoch, domain, drawn=None):
        """Module-by-module order of encoder/asr.py:103-119 for the rare configurations."""
        out = self.frontend(feat, feat_len)
        y, ylen = out[0], out[1]
        if hasattr(self, "normalize"):
            y, ylen = self.normalize(y, ylen, epoch=epoch, group_ids=domain)
        if self.training and hasattr(self, "specaug"):
            if drawn is None:
                y, ylen = self.specaug(y, ylen)
            else:
                warp, masks = drawn
                if warp is not None:
                    y, _, _ = ops.normalize(y, ylen, warp=warp)
                masks.fill_dev = ops.mean_all(y)
                y, _, _ = ops.normalize(y, ylen, dst=torch.empty_like(y), masks=masks)
        if len(out) == 4:
            return y, ylen, out[2], out[3]
        return y, ylen

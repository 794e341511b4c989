"""Fused frontend -> normalise -> SpecAugment chain.

The reference's ASREncoder calls the three modules one after the other (encoder/asr.py:103-119);
each is a drop-in here too, but calling them separately costs one extra read+write of the feature
tensor per module.  ``AcousticFrontend`` owns the three and issues the cheapest schedule:

* statistics known before the batch (eval, or global statistics frozen after ``max_epoch_num``):
  ONE kernel — normalisation and masks run in the epilogue of the fused frontend kernel and the
  feature tensor is written exactly once;
* utterance statistics: the frontend kernel writes the raw valid frames once and accumulates
  fixed-point sums, then one normalise(+mask) pass finishes in place — two launches, one library call;
* batch / global statistics while they are still being averaged: the same two launches with the
  (2·D+1)-float statistics exchange (one all-reduce under DDP) between them;
* a SpecAugment time warp is a 4-tap gather along time and cannot run in place: when a normalise pass exists
  it carries the warp (and the masks) out of place, otherwise the warp costs one extra pass.
"""
from typing import Optional

import torch

from ..host import ops
from ..host.ops import HostLengths
from .abs import Module
from .augment.specaug import SpecAugment
from .frontend._common import lengths_pair, squeeze_speech
from .frontend.speech2mel import Speech2MelSpec
from .norm.feat_norm import FeatureNormalization


class AcousticFrontend(Module):
    def module_init(self, frontend: dict, normalize=None, specaug: Optional[dict] = None, out_dtype=None):
        """``frontend`` / ``normalize`` / ``specaug`` take the same YAML blocks as ASREncoder
        (encoder/asr.py:44-68): frontend = {type:…, conf:{…}} or the conf dict itself;
        normalize = True | dict | None; specaug = dict | None.
        ``out_dtype="bfloat16"`` (extension): the last pass of the chain writes bfloat16 (B, F, D) — the layout and
        dtype Conv2dPrenet consumes under AMP (prenet/conv2d.py:220-276) — instead of fp32 followed by autocast's
        cast pass."""
        if out_dtype in ("bfloat16", "bf16", torch.bfloat16):
            self.out_dtype = torch.bfloat16
        elif out_dtype in (None, "float32", "fp32", torch.float32):
            self.out_dtype = None
        else:
            raise ValueError(f"out_dtype must be float32 or bfloat16, got {out_dtype!r}")
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

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_fast", None)  # (looked-up submodules: rebuilt on the next call)
        return d

    # ------------------------------------------------------------------------------------
    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor, epoch: int = None, domain=None):
        """feat (batch, T, 1) or (batch, T) fp32 CUDA waveforms, feat_len (batch,) int64 (CPU or CUDA).
        Returns (features (batch, F, n_mels), feature lengths (batch,) on the device)."""
        # (submodules live in nn.Module._modules: every `self.x` is a __getattr__ round trip — look them up once)
        fast = self.__dict__.get("_fast")
        if fast is None:
            fe = self.frontend
            fast = self.__dict__["_fast"] = (fe, fe.speech2linear, fe.linear2mel, getattr(self, "normalize", None),
                                             getattr(self, "specaug", None))
        fe, s2l, l2m, norm, aug = fast
        if not self.training:
            aug = None
        bf16 = self.__dict__.get("out_dtype") == torch.bfloat16
        if fe.delta_order is not None or fe.return_energy or s2l.pre_stft_norm is not None:
            if bf16:
                raise NotImplementedError("out_dtype=bfloat16 is available for the fused log-Mel chain only")
            return self._unfused(feat, feat_len, epoch, domain)
        if norm is not None and not (norm.mean_norm or norm.std_norm):
            norm = None

        if feat.dim() == 3 and feat.size(2) == 1 and feat.dtype == torch.float32:
            wav = feat  # (B, T, 1): the kernels take pointer, row pitch and (B, T) — no view needed
        else:
            wav = squeeze_speech(feat)
        if not wav.is_cuda:
            raise RuntimeError("speechain_b200 frontend: `speech` must be a CUDA tensor — there is no CPU path.")
        if wav.stride(1) != 1:
            wav = squeeze_speech(feat).contiguous()
        len_dev, len_host = lengths_pair(feat_len, wav.device)
        # feature lengths on the host: the shortest one now (the SpecAugment draws are bounded by it), the array only
        # if somebody asks for it (HostLengths.get on the returned lengths)
        flen_min = s2l._feat_len_of(int(len_host.min()))
        flen_host = None
        plan = s2l._plan(wav.device, preemph=True, mel=l2m.mel_conf())

        # SpecAugment draws need only host-side lengths, so they happen before any launch
        warp = masks = None
        if aug is not None:
            warp, masks = aug.draw_device(wav.size(0), fe.output_size, None, wav.device, minlen=flen_min)
            if masks is not None and not aug.feat_norm:
                if bf16:
                    raise NotImplementedError("out_dtype=bfloat16 is available for the fused log-Mel chain only")
                return self._unfused(feat, feat_len, epoch, domain, drawn=(warp, masks))
        # masks ride in a kernel epilogue; a time warp needs one out-of-place gather pass, which also
        # carries the normalisation and the masks when the statistics are only known after the first pass
        # (fp32 output: the library call carries the warp itself — frontend launch, then one warp + mask pass; a bf16
        #  output keeps the separate last pass, which also converts)
        in_call = warp is not None and not bf16
        fused_masks = masks if (warp is None or in_call) else None
        cwarp = warp if in_call else None
        warp_done = warp is None or in_call

        if norm is None:
            y, ylen = plan.chain(wav, len_dev, norm_mode=0, masks=fused_masks, len_host=len_host, warp=cwarp)
        else:
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
                mean = norm.get_buffer(f"{prefix}_mean") if norm.mean_norm else None
                std = norm.get_buffer(f"{prefix}_std") if norm.std_norm else None
                y, ylen = plan.chain(wav, len_dev, norm_mode=1, mean=mean, std=std, use_mean=norm.mean_norm,
                                     use_std=norm.std_norm, masks=fused_masks, len_host=len_host, warp=cwarp)
            elif not bf16 and plan.mel_logging and (kind == "utterance" or (kind == "batch" and not norm.training)):
                y, ylen = plan.chain(wav, len_dev, norm_mode=2, clamp=norm.clamp, use_mean=norm.mean_norm,
                                     use_std=norm.std_norm, masks=fused_masks, len_host=len_host, warp=cwarp)
            elif (not bf16 and plan.mel_logging and not isinstance(domain, torch.Tensor)
                  and norm.averaging(prefix, epoch)):
                # statistics still being averaged: two library calls with the (2 D + 1)-float exchange between them
                y, ylen = norm.average_step(plan, wav, len_dev, len_host, prefix, epoch, masks=masks, warp=warp)
                warp_done = True
            elif not bf16 and plan.mel_logging and kind == "batch" and norm.training:
                y, ylen = norm.batch_step(plan, wav, len_dev, len_host, masks=masks, warp=warp)
                warp_done = True
            else:
                # raw rows + sums, then ONE pass: statistics (exchanged across ranks if needed), warp, masks
                if wav.dim() == 3:
                    wav = wav.select(2, 0)
                out = plan.forward(wav, len_dev, want_stats=True, skip_pad=plan.mel_logging, len_host=len_host)
                y, ylen = out["mel"], out["feat_len"]
                y._sfe_stats = (y._version, out["stats"])
                flen_host = s2l._feat_len_host(len_host)
                HostLengths.remember(ylen, flen_host)
                y, ylen = norm.forward(y, ylen, group_ids=domain, epoch=epoch, masks=masks, inplace=True,
                                       src_has_pad=not plan.mel_logging, pad_value=plan.pad_mel, warp=warp,
                                       out_dtype=torch.bfloat16 if bf16 else None)
                warp_done = True
        HostLengths.remember(ylen, flen_host if flen_host is not None else (lambda: s2l._feat_len_host(len_host)))
        if not warp_done:
            y, _, _ = ops.normalize(y, ylen, warp=warp, masks=masks, out_bf16=bf16)
        elif bf16 and y.dtype != torch.bfloat16:
            # single-pass schedules end in the frontend kernel (fp32): one conversion pass of the normalise kernel
            y, _, _ = ops.normalize(y, ylen, use_mean=False, use_std=False, out_bf16=True)
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

"""Feature normalisation.  API mirror of speechain/module/norm/feat_norm.py
(`FeatureNormalization`): per-utterance statistics come from fp64 sums accumulated on the GPU
(by the fused frontend kernel when it produced ``feat``, else by one streaming pass), the
normalisation is one read + one write, and the DDP exchange is ONE all-reduce of a packed
(2·D+1)-float buffer instead of the reference's three all_gathers (feat_norm.py:146-266)."""
from typing import Dict, Optional

import numpy as np
import torch

from ...host import ops
from ...host.dist import all_reduce_packed
from ...host.ops import DeviceMasks, HostLengths
from ..abs import Module


class FeatureNormalization(Module):
    """Same constructor keys, defaults and lazily registered buffers as the reference
    (feat_norm.py:44-51, 455-531): ``global_mean/global_std/global_batch``, ``{gid}_mean/…``,
    ``aver_mean/aver_std``."""

    def module_init(self, norm_type: str = "global", mean_norm: bool = True, std_norm: bool = True,
                    clamp: float = 1e-10, max_epoch_num: int = 4):
        self.norm_type = norm_type
        self.mean_norm = mean_norm
        self.std_norm = std_norm
        self.clamp = float(clamp)
        self.max_epoch_num = max_epoch_num
        if self.input_size is not None:
            self.output_size = self.input_size
        self._batch_host: Dict[str, int] = {}  # host mirror of the {prefix}_batch counters (no sync per step)

    # ------------------------------------------------------------------------------------
    # statistics
    # ------------------------------------------------------------------------------------
    @staticmethod
    def _sums(feat: torch.Tensor, feat_len: torch.Tensor) -> torch.Tensor:
        """(B, D, 2) fp64 [sum, sum of squares] over the valid frames of each utterance."""
        stash = getattr(feat, "_sfe_stats", None)
        if stash is not None and stash[0] == feat._version:
            return stash[1]
        return ops.stats_accumulate(feat, feat_len)

    def _all_reduce(self, packed: torch.Tensor) -> torch.Tensor:
        return all_reduce_packed(packed) if self.distributed else packed

    def _counter(self, prefix: str) -> Optional[int]:
        if prefix in self._batch_host:
            return self._batch_host[prefix]
        if hasattr(self, f"{prefix}_batch"):  # restored from a checkpoint: one sync, once
            self._batch_host[prefix] = int(self.get_buffer(f"{prefix}_batch").item())
            return self._batch_host[prefix]
        return None

    def register_mean_std_batch(self, packed: torch.Tensor, D: int, prefix: str, epoch: Optional[int]):
        """Running average of feat_norm.py:455-508 from the packed [sum mean | sum std | N] sums."""
        dev = packed.device
        count = self._counter(prefix)
        if epoch is None or count is None:
            count = 1
        elif epoch <= self.max_epoch_num:
            count += 1
        self._batch_host[prefix] = count
        self.register_buffer(f"{prefix}_batch", torch.full((1,), count, dtype=torch.int64, device=dev))
        if not (epoch is None or epoch <= self.max_epoch_num):
            return  # frozen: like the reference, buffers that do not exist yet stay missing
        first_m = self.mean_norm and not hasattr(self, f"{prefix}_mean")
        first_s = self.std_norm and not hasattr(self, f"{prefix}_std")
        if first_m:
            self.register_buffer(f"{prefix}_mean", torch.empty(D, dtype=torch.float32, device=dev))
        if first_s:
            self.register_buffer(f"{prefix}_std", torch.empty(D, dtype=torch.float32, device=dev))
        weight = float(torch.tensor(1.0) / torch.tensor(count))  # fp32 1/count like `1 / LongTensor`
        run_mean = self.get_buffer(f"{prefix}_mean") if self.mean_norm else None
        run_std = self.get_buffer(f"{prefix}_std") if self.std_norm else None
        if first_m == first_s or run_mean is None or run_std is None:
            first = first_m if self.mean_norm else first_s
            ops.running_update(packed, D, weight, first, True, run_mean, run_std)
        else:  # one of the two buffers pre-exists (partial checkpoint)
            ops.running_update(packed, D, weight, first_m, True, run_mean, None)
            ops.running_update(packed, D, weight, first_s, True, None, run_std)

    def averaging(self, prefix: Optional[str], epoch: Optional[int]) -> bool:
        """True when a training step on `prefix` ("global" or a group name) updates the running statistics and both
        buffers are in the same state (both missing or both present) — the case AcousticFrontend runs as two library
        calls (`average_step`).  Everything else takes `forward`."""
        if not self.training or prefix is None or self.norm_type not in ("global", "group"):
            return False
        if not (epoch is None or epoch <= self.max_epoch_num):
            return False
        if self.mean_norm and self.std_norm:
            return hasattr(self, f"{prefix}_mean") == hasattr(self, f"{prefix}_std")
        return True

    def _advance(self, prefix: str, epoch: Optional[int], D: int, dev):
        """Host side of register_mean_std_batch (feat_norm.py:455-508) for a step that updates the running average:
        the counter, the fp32 weight 1/count, and the buffers (created on first use) the kernels write."""
        count = self._counter(prefix)
        count = 1 if (epoch is None or count is None) else count + 1
        self._batch_host[prefix] = count
        bufs = self._buffers
        counter = bufs.get(f"{prefix}_batch")
        if counter is None or counter.device != dev:
            self.register_buffer(f"{prefix}_batch", torch.empty((1,), dtype=torch.int64, device=dev))
            counter = bufs[f"{prefix}_batch"]
        run_mean = run_std = None
        first = False
        for name in (f"{prefix}_mean", f"{prefix}_std"):  # (a checkpoint loaded on another device)
            t = bufs.get(name)
            if t is not None and (t.device != dev or t.dtype != torch.float32 or not t.is_contiguous()):
                bufs[name] = t.to(dev, torch.float32).contiguous()
        if self.mean_norm:
            run_mean = bufs.get(f"{prefix}_mean")
            if run_mean is None:
                first = True
                self.register_buffer(f"{prefix}_mean", torch.empty(D, dtype=torch.float32, device=dev))
                run_mean = bufs[f"{prefix}_mean"]
        if self.std_norm:
            run_std = bufs.get(f"{prefix}_std")
            if run_std is None:
                first = True
                self.register_buffer(f"{prefix}_std", torch.empty(D, dtype=torch.float32, device=dev))
                run_std = bufs[f"{prefix}_std"]
        weight = float(np.float32(1.0) / np.float32(count))  # fp32 1/count like `1 / LongTensor`
        return count, weight, first, run_mean, run_std, counter

    def frozen(self, prefix: Optional[str], epoch: Optional[int]) -> bool:
        """True when a training step can no longer change the statistics of `prefix` (epoch > max_epoch_num) and the
        buffers it normalises with exist: the batch statistics — and their all-reduce — would be computed for nothing
        (the reference computes them and drops them, feat_norm.py:483-486)."""
        if prefix is None or epoch is None or epoch <= self.max_epoch_num:
            return False
        bufs = self._buffers
        return (f"{prefix}_batch" in bufs and (not self.mean_norm or f"{prefix}_mean" in bufs)
                and (not self.std_norm or f"{prefix}_std" in bufs))

    def average_step(self, plan, wav, len_dev, len_host, prefix: str, epoch: Optional[int], masks=None, warp=None):
        """The step of feat_norm.py:232-266, 455-508, 327-336 while the statistics are being averaged, straight from
        the waveforms: frontend launch with sums + packed batch statistics (one library call), the all-reduce of the
        2 D + 1 floats under DDP, then running average + counter + normalise [+ warp] [+ masks] (one more call)."""
        count, weight, first, run_mean, run_std, counter = self._advance(prefix, epoch, plan.n_mels, wav.device)
        mel, ylen, packed = plan.chain_sums(wav, len_dev, clamp=self.clamp, len_host=len_host)
        packed = self._all_reduce(packed)
        y = ops.running_normalize(packed, weight, first, run_mean, run_std, counter, count, mel, ylen,
                                  src_has_pad=False, pad_value=plan.pad_mel, masks=masks, warp=warp)
        return y, ylen

    def batch_step(self, plan, wav, len_dev, len_host, masks=None, warp=None):
        """``norm_type='batch'`` in training (feat_norm.py:232-299) straight from the waveforms, as the same two library
        calls as `average_step`: the batch statistics (all-reduced under DDP) land in a scratch pair of vectors
        instead of running buffers."""
        D = plan.n_mels
        mel, ylen, packed = plan.chain_sums(wav, len_dev, clamp=self.clamp, len_host=len_host)
        packed = self._all_reduce(packed)
        tmp = torch.empty((2, D), dtype=torch.float32, device=wav.device)
        y = ops.running_normalize(packed, 1.0, True, tmp[0] if self.mean_norm else None,
                                  tmp[1] if self.std_norm else None, None, 0, mel, ylen,
                                  src_has_pad=False, pad_value=plan.pad_mel, masks=masks, warp=warp)
        return y, ylen

    # ------------------------------------------------------------------------------------
    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor, group_ids=None, epoch: int = None,
                masks: Optional[DeviceMasks] = None, inplace: bool = False, src_has_pad: bool = True,
                pad_value: float = 0.0, warp=None, out_dtype=None):
        """feat (batch, length, channel) or (batch, length) fp32 CUDA; feat_len (batch,).
        ``masks`` / ``warp`` / ``inplace`` / ``src_has_pad`` are extensions used by AcousticFrontend to fuse
        SpecAugment (time warp gather + masking) into the same pass; ``out_dtype=torch.bfloat16`` makes that pass
        write bfloat16 (what an AMP prenet consumes); the reference signature is the first four."""
        if out_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError(f"out_dtype must be float32 or bfloat16, got {out_dtype}")
        bf16 = out_dtype == torch.bfloat16
        if self.norm_type == "group":
            assert group_ids is not None, (
                "You are using group-level feature normalization, but group_ids is not given. "
                "Please check 'data_cfg' in your configuration.")
        squeeze = False
        if feat.dim() == 2:
            feat, squeeze = feat.unsqueeze(-1), True
        elif feat.dim() != 3:
            raise RuntimeError(f"{self.__class__.__name__} only accepts the input vectors in the shape of "
                               f"(batch, length, channel) or (batch, length), but got shape={feat.shape}!")
        if not feat.is_cuda:
            raise RuntimeError("speechain_b200 FeatureNormalization: `feat` must be a CUDA tensor — no CPU path.")
        stash = getattr(feat, "_sfe_stats", None)
        if feat.dtype != torch.float32 or not feat.is_contiguous():
            feat = feat.float().contiguous()
            stash = None
        if stash is not None:
            feat._sfe_stats = stash
        if feat_len.dtype != torch.int64:
            feat_len = feat_len.long()
        if not feat_len.is_cuda:
            host = feat_len
            feat_len = feat_len.to(feat.device, non_blocking=True)
            HostLengths.remember(feat_len, host.numpy())
        B, _, D = feat.shape
        dst = feat if (inplace and warp is None) else None  # the warp gathers neighbouring rows: never in place
        common = dict(masks=masks, src_has_pad=src_has_pad, pad_value=pad_value, warp=warp, out_bf16=bf16)
        if bf16:
            dst = None  # (ops.normalize allocates the bfloat16 result)
        elif dst is None:
            dst = torch.empty_like(feat)
        if not (self.mean_norm or self.std_norm):
            out, _, _ = ops.normalize(feat, feat_len, dst=dst, **common)
            return (out.squeeze(-1) if squeeze else out), feat_len

        if self.norm_type == "utterance" or (self.norm_type == "batch" and not self.training):
            # feat_norm.py:137-139, 288-299: statistics are resolved inside the normalise kernel
            sums = self._sums(feat, feat_len)
            out, _, _ = ops.normalize(feat, feat_len, dst=dst, stats=sums, clamp=self.clamp, use_mean=self.mean_norm,
                                      use_std=self.std_norm, **common)
        elif self.norm_type == "group" and isinstance(group_ids, torch.Tensor):
            sums = self._sums(feat, feat_len) if self.training else None
            out = self._forward_group_tensor(feat, feat_len, dst, sums, group_ids, epoch, common)
        elif self.norm_type in ("batch", "global", "group"):
            if self.norm_type == "group":
                assert isinstance(group_ids, (str, int)), (
                    f"If all the utterances in a single batch belong to the same group, "
                    f"you should give group_ids as a string or integer. But got type(group_ids)={type(group_ids)}.")
            mean = std = None
            prefix = None if self.norm_type == "batch" else ("global" if self.norm_type == "global" else str(group_ids))
            training = self.training and not self.frozen(prefix, epoch)
            # (evaluation and frozen statistics need no sums: without the frontend kernel's they would cost a pass)
            sums = self._sums(feat, feat_len) if training else None
            if training and not bf16 and self.averaging(prefix, epoch):
                # statistics still being averaged: running average + counter + normalise in one library call
                _, _, packed = ops.stats_finalize(sums, feat_len, self.clamp, want_packed=True)
                packed = self._all_reduce(packed)
                count, weight, first, run_mean, run_std, counter = self._advance(prefix, epoch, D, feat.device)
                out = ops.running_normalize(packed, weight, first, run_mean, run_std, counter, count, feat, feat_len,
                                            src_has_pad=src_has_pad, pad_value=pad_value, masks=masks, warp=warp, dst=dst)
                return (out.squeeze(-1) if squeeze else out), feat_len
            if training:
                _, _, packed = ops.stats_finalize(sums, feat_len, self.clamp, want_packed=True)
                packed = self._all_reduce(packed)
                if self.norm_type == "batch":
                    mean, std = ops.running_update(packed, D, 1.0, False, False, None, None, want_batch=True)
                    mean = mean if self.mean_norm else None
                    std = std if self.std_norm else None
            if self.norm_type != "batch":
                if training:
                    self.register_mean_std_batch(packed, D, prefix, epoch)
                mean = self.get_buffer(f"{prefix}_mean") if self.mean_norm else None
                std = self.get_buffer(f"{prefix}_std") if self.std_norm else None
            out, _, _ = ops.normalize(feat, feat_len, dst=dst, mean=mean, std=std, per_utt=False, **common)
        else:
            raise AssertionError(f"norm_type can only be one of 'utterance', 'batch', 'group', 'global', "
                                 f"but got norm_type={self.norm_type}!")
        return (out.squeeze(-1) if squeeze else out), feat_len

    # ------------------------------------------------------------------------------------
    # group-level normalisation with tensor ids (speaker-wise statistics, feat_norm.py:153-230)
    # ------------------------------------------------------------------------------------
    def _forward_group_tensor(self, feat, feat_len, dst, sums, group_ids, epoch, common):
        B, _, D = feat.shape
        ids_host = [int(g) for g in group_ids.tolist()]
        if self.training:
            mean, std, _ = ops.stats_finalize(sums, feat_len, self.clamp, want_packed=False)
            if self.distributed and torch.distributed.is_available() and torch.distributed.is_initialized():
                mean, std, ids_host = self._gather_group(mean, std, ids_host)
            # bookkeeping on (n_groups, D)-sized tensors; the per-group mean of per-utterance stats
            order = []
            for g in ids_host:
                if g not in order:
                    order.append(g)
            for g in order:
                sel = torch.tensor([i for i, x in enumerate(ids_host) if x == g], device=feat.device)
                gm = mean.index_select(0, sel).mean(dim=0) if self.mean_norm else None
                gs = std.index_select(0, sel).mean(dim=0) if self.std_norm else None
                self._register_group(str(g), gm, gs, epoch)
            self.update_aver_mean_std(epoch)
            ids_host = [int(g) for g in group_ids.tolist()]  # back to the local utterances
        rows_m, rows_s = [], []
        for g in ids_host:
            if self.mean_norm:
                rows_m.append(self.get_buffer(f"{g}_mean") if hasattr(self, f"{g}_mean") else self.get_buffer("aver_mean"))
            if self.std_norm:
                rows_s.append(self.get_buffer(f"{g}_std") if hasattr(self, f"{g}_std") else self.get_buffer("aver_std"))
        m = torch.stack(rows_m).contiguous() if rows_m else None
        s = torch.stack(rows_s).contiguous() if rows_s else None
        out, _, _ = ops.normalize(feat, feat_len, dst=dst, mean=m, std=s, per_utt=True, **common)
        return out

    def _gather_group(self, mean, std, ids_host):
        """all_gather of per-utterance statistics + ids (feat_norm.py:157-171), padded to the
        largest local batch."""
        world = torch.distributed.get_world_size()
        dev = mean.device
        n = torch.tensor([len(ids_host)], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        torch.distributed.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        mx = max(sizes)
        D = mean.size(1)
        pack = torch.zeros((mx, 2 * D + 1), dtype=torch.float32, device=dev)
        pack[: len(ids_host), :D] = mean
        pack[: len(ids_host), D:2 * D] = std
        pack[: len(ids_host), 2 * D] = torch.tensor(ids_host, dtype=torch.float32, device=dev)
        allp = [torch.zeros_like(pack) for _ in range(world)]
        torch.distributed.all_gather(allp, pack)
        cat = torch.cat([p[:k] for p, k in zip(allp, sizes)])
        return cat[:, :D].contiguous(), cat[:, D:2 * D].contiguous(), [int(v) for v in cat[:, 2 * D].tolist()]

    def _register_group(self, prefix: str, cur_mean, cur_std, epoch):
        dev = (cur_mean if cur_mean is not None else cur_std).device
        count = self._counter(prefix)
        if epoch is None or count is None:
            count = 1
        elif epoch <= self.max_epoch_num:
            count += 1
        self._batch_host[prefix] = count
        self.register_buffer(f"{prefix}_batch", torch.full((1,), count, dtype=torch.int64, device=dev))
        if epoch is None or epoch <= self.max_epoch_num:
            w = torch.tensor(1.0, device=dev) / count
            for name, cur, on in ((f"{prefix}_mean", cur_mean, self.mean_norm), (f"{prefix}_std", cur_std, self.std_norm)):
                if not on:
                    continue
                if not hasattr(self, name):
                    self.register_buffer(name, cur)
                else:
                    self.register_buffer(name, w * cur + (1 - w) * self.get_buffer(name))

    def update_aver_mean_std(self, epoch: int):
        """feat_norm.py:510-531 (iterates *all* buffers ending in _mean/_std, aver_* included)."""
        if epoch is None or epoch <= self.max_epoch_num:
            n_m = n_s = 0
            a_m = a_s = None
            for name, buff in self.named_buffers():
                if name.endswith("_mean"):
                    n_m += 1
                    a_m = buff.clone() if a_m is None else a_m + buff
                elif name.endswith("_std"):
                    n_s += 1
                    a_s = buff.clone() if a_s is None else a_s + buff
            if a_m is not None:
                self.register_buffer("aver_mean", a_m / n_m)
            if a_s is not None:
                self.register_buffer("aver_std", a_s / n_s)

    # ------------------------------------------------------------------------------------
    def recover(self, feat: torch.Tensor, group_ids=None):
        """Inverse of the global / group normalisation (feat_norm.py:533-607); a (D,)-broadcast on
        tiny statistics, not part of the accelerated forward path."""
        assert self.norm_type not in ["utterance", "batch"], \
            "If norm_type is either 'utterance' or 'batch', the normalized features cannot be recovered."
        if self.norm_type == "global" or (self.norm_type == "group" and isinstance(group_ids, (str, int))):
            prefix = "global" if self.norm_type == "global" else str(group_ids)
            feat = feat * self.get_buffer(f"{prefix}_std") if self.std_norm else feat
            feat = feat + self.get_buffer(f"{prefix}_mean") if self.mean_norm else feat
        elif self.norm_type == "group" and isinstance(group_ids, torch.Tensor):
            ids = [int(g) for g in group_ids.tolist()]
            if self.std_norm:
                feat = feat * torch.stack([self.get_buffer(f"{g}_std") if hasattr(self, f"{g}_std")
                                           else self.get_buffer("aver_std") for g in ids]).unsqueeze(1)
            if self.mean_norm:
                feat = feat + torch.stack([self.get_buffer(f"{g}_mean") if hasattr(self, f"{g}_mean")
                                           else self.get_buffer("aver_mean") for g in ids]).unsqueeze(1)
        elif self.norm_type == "group" and group_ids is None:
            feat = feat * self.get_buffer("aver_std").expand(1, 1, -1) if self.std_norm else feat
            feat = feat + self.get_buffer("aver_mean").expand(1, 1, -1) if self.mean_norm else feat
        else:
            raise RuntimeError
        return feat

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """Lazily register every buffer ending in _batch/_std/_mean found in the checkpoint
        (feat_norm.py:609-632)."""
        self._batch_host.clear()
        for key in state_dict.keys():
            if key.startswith(prefix):
                name = key[len(prefix):].split(".", 1)[0]
                if "_" in name and name.split("_")[-1] in ["batch", "std", "mean"]:
                    self.register_buffer(name, state_dict[key])
                else:
                    unexpected_keys.append(key)

    def __getstate__(self):
        return self.__dict__.copy()

    def extra_repr(self) -> str:
        return f"norm_type={self.norm_type}, mean_norm={self.mean_norm}, std_norm={self.std_norm}"

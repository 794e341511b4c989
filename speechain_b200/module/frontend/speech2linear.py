"""Speech -> linear spectrogram.  API mirror of speechain/module/frontend/speech2linear.py
(`Speech2LinearSpec`); the arithmetic is the fused sm_100a kernel behind ``sfe_frontend_forward``."""
from typing import Optional

import torch

from ...host import ops
from ...host.ops import FrontendPlan, HostLengths
from ..abs import Module
from ._common import ConstHolder, as_float, lengths_pair, make_window, squeeze_speech, to_samples


class Speech2LinearSpec(Module):
    """Waveform -> linear (energy or magnitude) spectrogram, optionally log-compressed.

    Same constructor keys and defaults as the reference (speech2linear.py:20-37).  Differences,
    all deliberate (SURVEY.md §8b "Ownership"): the caller's ``speech_len`` is not modified in
    place (speech2linear.py:202 does), the input waveform is never written to, and a 2-D
    ``(batch, time)`` input also works with pre-emphasis.
    """

    def module_init(self, hop_length, win_length, sr: int = 16000, n_fft: int = None, preemphasis: float = None,
                    pre_stft_norm: str = None, window: str = "hann", center: bool = True, normalized: bool = False,
                    onesided: bool = True, mag_spec: bool = False, return_energy: bool = False, clamp: float = 1e-10,
                    logging: bool = False, log_base: float = None):
        hop_length = to_samples(hop_length, sr)
        win_length = to_samples(win_length, sr)
        if n_fft is None:
            n_fft = win_length  # speech2linear.py:93-94
        if pre_stft_norm not in (None, "mean_std", "min_max"):
            raise ValueError(f"pre_stft_norm must be None, 'mean_std' or 'min_max', got {pre_stft_norm!r}")
        # (two-sided: all n_fft bins of torch.stft(onesided=False); LinearSpec2MelSpec cannot consume those)
        self.onesided = bool(onesided)
        self.output_size = n_fft // 2 + 1 if onesided else n_fft
        self.win_length, self.hop_length, self.n_fft, self.sr = win_length, hop_length, n_fft, sr
        self.preemphasis = as_float(preemphasis, "preemphasis")
        if self.preemphasis is not None:
            self.preemph_filter = ConstHolder(torch.tensor([-self.preemphasis, 1.0]).reshape(1, 1, 2))
        self.pre_stft_norm = pre_stft_norm
        self.stft_config = dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                                center=center, normalized=normalized, onesided=onesided, return_complex=True)
        self.mag_spec = mag_spec
        self.return_energy = return_energy
        self.clamp = as_float(clamp, "clamp")
        self.logging = logging
        self.log_base = as_float(log_base, "log_base")
        self._plans = {}

    # ------------------------------------------------------------------------------------
    def _plan(self, device, preemph: bool = True, mel: Optional[dict] = None, plain_window: bool = False,
              raw_power: bool = False) -> FrontendPlan:
        # (a checkpoint load copies into the filterbank buffer in place: the version counter catches that)
        key = (device, preemph, None if mel is None else (id(mel["weight"]), mel["weight"]._version), plain_window, raw_power)
        plan = self._plans.get(key)
        if plan is None:
            c = self.stft_config
            # (Griffin-Lim in the reference always runs un-normalised STFTs, whatever `normalized` says)
            win = make_window(c["window"], self.win_length, self.n_fft, c["normalized"] and not plain_window)
            kw = dict(n_fft=self.n_fft, win_length=self.win_length, hop_length=self.hop_length, center=c["center"],
                      preemphasis=self.preemphasis if preemph else None, mag_spec=self.mag_spec and not raw_power,
                      lin_logging=self.logging and not raw_power, lin_clamp=self.clamp, lin_log_base=self.log_base, window=win,
                      device=device)
            if mel is not None:
                kw.update(n_mels=mel["n_mels"], mel_logging=mel["logging"], mel_clamp=mel["clamp"],
                          mel_log_base=mel["log_base"], mel_weight=mel["weight"])
            plan = FrontendPlan(**kw)
            self._plans[key] = plan
        return plan

    def _prepare(self, speech: torch.Tensor, speech_len: torch.Tensor):
        """Shared by Speech2LinearSpec and Speech2MelSpec: shape checks, optional waveform
        pre-normalisation, device lengths.  Returns (wav (B,T), len_dev, len_host, preemph_in_kernel)."""
        wav = squeeze_speech(speech)
        if not wav.is_cuda:
            raise RuntimeError("speechain_b200 frontend: `speech` must be a CUDA tensor — there is no CPU path.")
        len_dev, len_host = lengths_pair(speech_len, wav.device)
        in_kernel = True
        if self.pre_stft_norm is not None:
            # Rare path (no shipped recipe sets it): the statistics need the pre-emphasised signal including padding
            # (speech2linear.py:177-187): one reduction + one apply kernel write the normalised rows, and the frontend
            # kernel then runs without pre-emphasis over whole rows.
            in_kernel = False
            if self.pre_stft_norm not in ("mean_std", "min_max"):
                raise ValueError(f"pre_stft_norm must be 'mean_std' or 'min_max', got {self.pre_stft_norm!r}")
            wav = ops.prestft_norm(wav, len_dev, preemphasis=self.preemphasis, mode=self.pre_stft_norm)
        return wav, len_dev, len_host, in_kernel

    def _kernel_lengths(self, wav, len_dev, in_kernel):
        """After a waveform pre-normalisation the padding is no longer silence (min_max maps 0 to a constant),
        and the reference's STFT sees it: let the kernel read whole rows.  (The kernel otherwise defines
        samples beyond a length as zero.)"""
        return len_dev if in_kernel else torch.full_like(len_dev, wav.size(1))

    def _fix_padded_rows(self, out, len_dev, in_kernel, plan, raw_power: bool = False):
        """Rare pre_stft_norm path: the kernel ran with full-row lengths, so rows t >= feat_len still hold
        spectra of the padding; reset them like speech2linear.py:226-233 / linear2mel.py:177-187 do."""
        if in_kernel:
            return
        # rows beyond each length: the normalise kernel synthesises them (identity statistics, pad value)
        feat_len = torch.from_numpy(self._feat_len_host(HostLengths.get(len_dev))).to(len_dev.device)
        if out.get("mel") is not None:
            ops.normalize(out["mel"], feat_len, use_mean=False, use_std=False, src_has_pad=False, pad_value=plan.pad_mel)
        if out.get("lin") is not None:
            import math
            pad_lin = 0.0
            if self.logging and not raw_power:
                pad_lin = math.log(self.clamp) / (math.log(self.log_base) if self.log_base is not None else 1.0)
            ops.normalize(out["lin"], feat_len, use_mean=False, use_std=False, src_has_pad=False, pad_value=pad_lin)
        if out.get("energy") is not None:
            ops.normalize(out["energy"].unsqueeze(-1), feat_len, use_mean=False, use_std=False, src_has_pad=False,
                          pad_value=0.0)
        out["feat_len"] = feat_len

    def _feat_len_host(self, len_host):
        """numpy int64 feature lengths from numpy sample lengths (speech2linear.py:201-212)."""
        # (x // hop + 1 == (x + hop) // hop: two array operations instead of four)
        c = (2 * (self.n_fft // 2) if self.stft_config["center"] else 0) - self.n_fft + self.hop_length
        return (len_host + c) // self.hop_length

    def _feat_len_of(self, n: int) -> int:
        """the same for one length (a Python int)"""
        c = (2 * (self.n_fft // 2) if self.stft_config["center"] else 0) - self.n_fft
        return (n + c) // self.hop_length + 1

    def __getstate__(self):
        # compiled plans hold raw library handles: never pickle / deepcopy them
        d = self.__dict__.copy()
        d["_plans"] = {}
        return d

    def forward(self, speech: torch.Tensor, speech_len: torch.Tensor):
        """speech (batch, T, 1) or (batch, T) fp32 CUDA; speech_len (batch,) int64.
        Returns (linear_spec (batch, F, n_fft//2+1), feat_len) or, with ``return_energy``,
        (linear_spec, feat_len, energy (batch, F), energy_len)."""
        wav, len_dev, len_host, in_kernel = self._prepare(speech, speech_len)
        two = not getattr(self, "onesided", True)
        # (two-sided: the kernel delivers the plain power spectrum, one more pass mirrors it, sums the energy over all
        #  n_fft bins and applies the point-wise tail)
        plan = self._plan(wav.device, preemph=in_kernel, raw_power=two)
        out = plan.forward(wav, self._kernel_lengths(wav, len_dev, in_kernel), want_mel=False, want_lin=True,
                           want_energy=self.return_energy and not two)
        self._fix_padded_rows(out, len_dev, in_kernel, plan, raw_power=two)
        feat_len = out["feat_len"]
        HostLengths.remember(feat_len, self._feat_len_host(len_host))
        if two:
            out["lin"], out["energy"] = ops.twosided(out["lin"], feat_len, self.n_fft, mag_spec=self.mag_spec,
                                                     logging=self.logging, clamp=self.clamp, log_base=self.log_base,
                                                     want_energy=self.return_energy)
        if self.return_energy:
            return out["lin"], feat_len, out["energy"], feat_len
        return out["lin"], feat_len

    # torchaudio.transforms.GriffinLim defaults the reference relies on (speech2linear.py:281-292)
    GL_N_ITER, GL_MOMENTUM = 32, 0.99

    def recover(self, feat: torch.Tensor, feat_len: torch.Tensor, inv_preemph_winlen: int = 100):
        """Linear spectrogram (batch, F, n_fft//2+1) -> (waveform, wav_len) by Griffin-Lim, inverse pre-emphasis over a
        sliding window and silencing of the padding (speech2linear.py:252-351).  Returns the reference's shapes:
        (batch, L, 1) when a pre-emphasis is undone, (batch, L) otherwise, with L = hop_length * (F - 1).

        The initial phase is drawn exactly like torchaudio does on the CPU — ``torch.rand`` of the (batch, bins,
        frames) complex tensor on the global CPU generator — so a seeded call reproduces the reference's CPU result."""
        if not feat.is_cuda:
            raise RuntimeError("speechain_b200 frontend: `feat` must be a CUDA tensor — there is no CPU path.")
        assert self.pre_stft_norm is None, "waveform pre-stft normalization cannot be recovered for TTS synthesis."
        B, F, K = feat.shape
        if K != self.n_fft // 2 + 1:
            raise ValueError(f"recover() needs the one-sided spectrogram (n_fft // 2 + 1 = {self.n_fft // 2 + 1} bins), "
                             f"got {K} (torchaudio's GriffinLim rejects it the same way)")
        len_host = HostLengths.get(feat_len)
        assert int(len_host.max()) == F, "Something wrong happens when calculating the length of synthetic utterances."
        len_dev = feat_len if feat_len.is_cuda else feat_len.to(feat.device, non_blocking=True)
        angles = torch.rand((B, K, F), dtype=torch.complex64).transpose(1, 2).contiguous()
        plan = self._plan(feat.device, preemph=True, plain_window=bool(self.stft_config["normalized"]))
        wav, wav_len = plan.griffinlim(feat.float(), len_dev.long(), n_iter=self.GL_N_ITER, momentum=self.GL_MOMENTUM,
                                       init_angles=angles.pin_memory().to(feat.device, non_blocking=True),
                                       inv_preemph_taps=inv_preemph_winlen if self.preemphasis is not None else 0)
        HostLengths.remember(wav_len, (len_host - 1) * self.hop_length)
        if self.preemphasis is not None:
            wav = wav.unsqueeze(-1)
        return wav, wav_len

    def get_sample_rate(self):
        return self.sr

    def __repr__(self) -> str:
        s = (f"{self.__class__.__name__}(\n"
             f"win_length={self.win_length}, hop_length={self.hop_length}, n_fft={self.n_fft}, mag_spec={self.mag_spec}, ")
        if self.preemphasis is not None:
            s += f"preemphasis={self.preemphasis}, "
        if self.pre_stft_norm is not None:
            s += f"pre_stft_norm={self.pre_stft_norm}, "
        return s + "\n)"

"""Speech -> (log-)mel.  API mirror of speechain/module/frontend/speech2mel.py (`Speech2MelSpec`):
one fused kernel launch instead of the reference's ~20 library launches and 3·B host syncs."""
import torch

from ...host.ops import HostLengths
from ..abs import Module
from ._common import to_samples
from .delta_feat import DeltaFeature
from .linear2mel import LinearSpec2MelSpec
from .speech2linear import Speech2LinearSpec


class Speech2MelSpec(Module):
    """Same constructor keys and defaults as the reference (speech2mel.py:24-48); children keep
    the names ``speech2linear`` / ``linear2mel`` / ``delta`` so state-dict keys match."""

    def module_init(self, n_mels: int, hop_length, win_length, n_fft: int = None, sr: int = 16000,
                    preemphasis: float = None, pre_stft_norm: str = None, window: str = "hann", center: bool = True,
                    normalized: bool = False, onesided: bool = True, mag_spec: bool = False, return_energy: bool = False,
                    fmin: float = 0.0, fmax: float = None, clamp: float = 1e-10, logging: bool = True,
                    log_base: float = 10.0, mel_scale: str = "slaney", mel_norm: bool = True, delta_order: int = None,
                    delta_N: int = 2):
        hop_length = to_samples(hop_length, sr)
        win_length = to_samples(win_length, sr)
        if n_fft is None:
            n_fft = win_length
        self.output_size = n_mels if delta_order is None else n_mels * (delta_order + 1)
        self.return_energy = return_energy
        self.speech2linear = Speech2LinearSpec(
            n_fft=n_fft, sr=sr, hop_length=hop_length, win_length=win_length, preemphasis=preemphasis,
            pre_stft_norm=pre_stft_norm, window=window, center=center, normalized=normalized, onesided=onesided,
            mag_spec=mag_spec, return_energy=return_energy, logging=False)  # speech2mel.py:136
        self.linear2mel = LinearSpec2MelSpec(
            sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax, clamp=clamp, logging=logging, log_base=log_base,
            mel_scale=mel_scale, mel_norm=mel_norm, mag_spec=mag_spec)
        self.delta_order = delta_order
        if delta_order is not None:
            self.delta_N = delta_N
            self.delta = DeltaFeature(delta_order=delta_order, delta_N=delta_N)
        # when True the kernel also accumulates per-utterance sums that a following
        # FeatureNormalization picks up instead of re-reading the features
        self.collect_stats = True

    def run(self, speech: torch.Tensor, speech_len: torch.Tensor, **kernel_kw):
        """The fused launch with extra epilogue options (used by AcousticFrontend).  Returns the
        raw dict of ``FrontendPlan.forward`` plus the host feature lengths."""
        s2l = self.speech2linear
        wav, len_dev, len_host, in_kernel = s2l._prepare(speech, speech_len)
        plan = s2l._plan(wav.device, preemph=in_kernel, mel=self.linear2mel.mel_conf())
        if not in_kernel and kernel_kw.get("want_stats"):
            kernel_kw = dict(kernel_kw, want_stats=False)  # sums of the padded rows would be wrong on this path
        out = plan.forward(wav, s2l._kernel_lengths(wav, len_dev, in_kernel), want_mel=True,
                           want_energy=self.return_energy, len_host=len_host if in_kernel else None, **kernel_kw)
        s2l._fix_padded_rows(out, len_dev, in_kernel, plan)
        out["feat_len_host"] = s2l._feat_len_host(len_host)
        out["pad_value"] = plan.pad_mel
        HostLengths.remember(out["feat_len"], out["feat_len_host"])
        return out

    def forward(self, speech: torch.Tensor, speech_len: torch.Tensor):
        """speech (batch, T, 1) or (batch, T) fp32 CUDA; speech_len (batch,) int64.
        Returns (feat (batch, F, output_size), feat_len[, energy (batch, F), energy_len])."""
        stats = self.collect_stats and self.delta_order is None
        out = self.run(speech, speech_len, want_stats=stats)
        feat, feat_len = out["mel"], out["feat_len"]
        if self.delta_order is not None:
            feat, feat_len = self.delta(feat, feat_len)
        elif stats and out.get("stats") is not None:
            # side channel for FeatureNormalization (valid while `feat` is not modified in place)
            feat._sfe_stats = (feat._version, out["stats"])
        if self.return_energy:
            return feat, feat_len, out["energy"], feat_len
        return feat, feat_len

    def recover(self, feat: torch.Tensor, feat_len: torch.Tensor):
        assert self.delta_order is None
        feat = self.linear2mel.recover(feat, feat_len)
        return self.speech2linear.recover(feat, feat_len)

    def get_sample_rate(self):
        return self.speech2linear.sr

    def __repr__(self):
        s = f"{self.__class__.__name__}(\n" + str(self.speech2linear) + "\n" + str(self.linear2mel)
        if self.delta_order is not None:
            s += "\n" + str(self.delta)
        return s + "\n)"

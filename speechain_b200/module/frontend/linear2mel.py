"""Linear spectrogram -> (log-)mel.  API mirror of speechain/module/frontend/linear2mel.py."""
import torch

from ...host.ops import FrontendPlan, HostLengths, mel2linear
from ..abs import Module
from ._common import ConstHolder, as_float, mel_filterbank


class LinearSpec2MelSpec(Module):
    """Same constructor keys and defaults as the reference (linear2mel.py:20-33).  The dense
    ``nn.Linear`` of the reference (97.6 % zeros) becomes a banded filter in the kernel; the full
    matrix is still kept as the ``mel_fbanks.weight`` buffer for checkpoint compatibility."""

    def module_init(self, n_fft: int, n_mels: int, sr: int = 16000, fmin: float = 0.0, fmax: float = None,
                    clamp: float = 1e-10, logging: bool = True, log_base: float = 10.0, mel_scale: str = "slaney",
                    mel_norm: bool = True, mag_spec: bool = False):
        self.sr = sr
        self.n_fft = n_fft
        self.stft_dim = n_fft // 2 + 1
        self.n_mels = n_mels
        self.fmin = fmin
        self.fmax = fmax if fmax is not None else sr // 2
        assert self.fmax > self.fmin, f"fmax must be larger than fmin, but got fmax={self.fmax} and fmin={self.fmin}!"
        assert mel_scale in ["htk", "slaney"], f"mel_scale must be either 'htk' or 'slaney', but got mel_scale={mel_scale}"
        self.mel_scale = mel_scale
        self.mel_norm = mel_norm
        weight = mel_filterbank(sr, n_fft, n_mels, float(self.fmin), float(self.fmax), mel_scale, mel_norm)
        self.mel_fbanks = ConstHolder(weight)
        self.input_size_expected = self.stft_dim
        self.output_size = n_mels
        self.clamp = as_float(clamp, "clamp")
        self.logging = logging
        self.log_base = as_float(log_base, "log_base")
        self.mag_spec = mag_spec  # stored, unused in forward — like linear2mel.py:161
        self._plans = {}

    def mel_conf(self) -> dict:
        return dict(n_mels=self.n_mels, logging=self.logging, clamp=self.clamp, log_base=self.log_base,
                    weight=self.mel_fbanks.weight)

    def _plan(self, device) -> FrontendPlan:
        key = (str(device), self.mel_fbanks.weight._version)
        plan = self._plans.get(key)
        if plan is None:
            plan = FrontendPlan(n_fft=self.n_fft, win_length=self.n_fft, hop_length=1, window=torch.ones(self.n_fft),
                                n_mels=self.n_mels, mel_logging=self.logging, mel_clamp=self.clamp,
                                mel_log_base=self.log_base, mel_weight=self.mel_fbanks.weight, device=device)
            self._plans[key] = plan
        return plan

    def __getstate__(self):
        # compiled plans hold raw library handles: never pickle / deepcopy them
        d = self.__dict__.copy()
        d["_plans"] = {}
        d.pop("_pinv", None)
        return d

    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor):
        """feat (batch, F, n_fft//2+1) fp32 CUDA -> ((log-)mel (batch, F, n_mels), feat_len)."""
        if feat.size(-1) != self.stft_dim:
            raise RuntimeError(f"expected {self.stft_dim} linear bins, got feat.shape={tuple(feat.shape)}")
        if not feat.is_cuda:
            raise RuntimeError("speechain_b200 frontend: `feat` must be a CUDA tensor — there is no CPU path.")
        return self._plan(feat.device).linear2mel(feat.float()), feat_len

    def _pinv_t(self, device) -> torch.Tensor:
        """pinv(W)^T (n_mels, n_fft//2+1) on `device`.  For the full-row-rank filterbank this is the
        minimum-norm solution operator that ``torch.linalg.lstsq(W, ·, driver="gels")`` applies
        (linear2mel.py:209-214); computed once per weight version in fp64 on the host."""
        w = self.mel_fbanks.weight
        key = (str(device), w._version)
        hit = self._pinv.get(key) if hasattr(self, "_pinv") else None
        if hit is None:
            self._pinv = {key: torch.linalg.pinv(w.detach().to("cpu", torch.float64)).T.contiguous().float().to(device)}
            hit = self._pinv[key]
        return hit

    def recover(self, feat: torch.Tensor, feat_len: torch.Tensor):
        """(log-)mel (batch, F, n_mels) -> linear spectrogram (batch, F, n_fft//2+1), clamped at 1e-10; rows
        beyond the length of every utterance shorter than the longest are silenced first (linear2mel.py:191-224)."""
        if not feat.is_cuda:
            raise RuntimeError("speechain_b200 frontend: `feat` must be a CUDA tensor — there is no CPU path.")
        if feat.size(-1) != self.n_mels:
            raise RuntimeError(f"expected {self.n_mels} mel bins, got feat.shape={tuple(feat.shape)}")
        len_host = HostLengths.get(feat_len)
        len_dev = feat_len if feat_len.is_cuda else feat_len.to(feat.device, non_blocking=True)
        mode = 0 if not self.logging else (1 if self.log_base is None else 2)
        return mel2linear(feat.float(), len_dev.long(), self._pinv_t(feat.device), log_mode=mode,
                          log_base=self.log_base or 0.0, max_len=int(len_host.max()) if len(len_host) else 0)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(\n"
                f"stft_dim={self.stft_dim}, n_mels={self.n_mels}, fmin={self.fmin}, fmax={self.fmax}, "
                f"mel_scale={self.mel_scale}, mel_norm={self.mel_norm}\n)")

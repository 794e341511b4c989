"""The reference's own modules (staged by oracle/build_ref.py under oracle/_ref) behind the function interface of
oracle/ref_port_torch.py, so that bench.py's CPU arm can time either.

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing in ``speechain_b200/`` may import this.

Each call builds (and caches) the reference module for the given configuration and runs its ``forward`` —
Speech2LinearSpec.forward (speech2linear.py:142-250), LinearSpec2MelSpec.forward (linear2mel.py:163-189),
FeatureNormalization.forward (feat_norm.py:78-338), SpecAugment.forward (specaug.py:73-206) — unmodified.  The one
exception is the same as in tests/golden/make_golden.py: ``FeatureNormalization.register_mean_std_batch`` calls
``.cuda()`` on a CPU tensor (feat_norm.py:475) and is overridden with the same arithmetic minus that call.
"""
import os
import sys
import warnings

import torch

_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available() -> bool:
    return os.path.exists(os.path.join(_ROOT, "MANIFEST"))


def _load():
    if not _mods:
        if _ROOT not in sys.path:
            sys.path.insert(0, _ROOT)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)  # speech2mel.py:49 has an invalid escape in a docstring
            from speechain.module.augment.specaug import SpecAugment
            from speechain.module.frontend.linear2mel import LinearSpec2MelSpec
            from speechain.module.frontend.speech2linear import Speech2LinearSpec
            from speechain.module.norm.feat_norm import FeatureNormalization

        class CpuFeatNorm(FeatureNormalization):
            def register_mean_std_batch(self, curr_aver_mean, curr_aver_std, prefix, epoch):
                if epoch is None or not hasattr(self, f"{prefix}_batch"):
                    self.register_buffer(f"{prefix}_batch", torch.LongTensor([1]))
                elif epoch <= self.max_epoch_num:
                    self.register_buffer(f"{prefix}_batch", self.get_buffer(f"{prefix}_batch") + 1)
                if epoch is None or epoch <= self.max_epoch_num:
                    w = 1 / self.get_buffer(f"{prefix}_batch")
                    for name, cur, on in (("mean", curr_aver_mean, self.mean_norm), ("std", curr_aver_std, self.std_norm)):
                        if not on:
                            continue
                        if not hasattr(self, f"{prefix}_{name}"):
                            self.register_buffer(f"{prefix}_{name}", cur)
                        else:
                            self.register_buffer(f"{prefix}_{name}", w * cur + (1 - w) * self.get_buffer(f"{prefix}_{name}"))

        _mods.update(SpecAugment=SpecAugment, LinearSpec2MelSpec=LinearSpec2MelSpec, Speech2LinearSpec=Speech2LinearSpec,
                     FeatureNormalization=CpuFeatNorm)
    return _mods


_cache = {}


def _get(kind, **conf):
    key = (kind, repr(sorted(conf.items())))
    if key not in _cache:
        _cache[key] = _load()[kind](**conf)
    return _cache[key]


def mel_weight(sr, n_fft, n_mels, fmin=0.0, fmax=None, mel_scale="slaney", mel_norm=True):
    """Here the "weight" is the configured LinearSpec2MelSpec module itself (linear2mel.py:117-161)."""
    return dict(n_fft=n_fft, n_mels=n_mels, sr=sr, fmin=fmin, fmax=fmax, mel_scale=mel_scale, mel_norm=mel_norm)


def speech2linear(speech, speech_len, *, n_fft, hop, win, preemphasis=None, mag_spec=False, return_energy=False, **_):
    m = _get("Speech2LinearSpec", hop_length=hop, win_length=win, n_fft=n_fft, preemphasis=preemphasis, mag_spec=mag_spec,
             return_energy=return_energy)
    out = m(speech, speech_len)
    return (out[0], out[1], out[2]) if return_energy else (out[0], out[1], None)


def linear2mel(spec, weight, *, clamp=1e-10, logging=True, log_base=10.0):
    m = _get("LinearSpec2MelSpec", clamp=clamp, logging=logging, log_base=log_base, **weight)
    return m(spec, torch.full((spec.size(0),), spec.size(1), dtype=torch.long))[0]


def norm_utterance(feat, feat_len, clamp=1e-10):
    return _get("FeatureNormalization", norm_type="utterance", clamp=clamp)(feat, feat_len)[0]


def norm_global_train(feat, feat_len, state: dict, epoch, max_epoch_num=4, clamp=1e-10):
    m = state.setdefault("module", _load()["FeatureNormalization"](norm_type="global", clamp=clamp, max_epoch_num=max_epoch_num))
    m.train()
    return m(feat, feat_len, epoch=epoch)[0]


def specaug(feat, feat_len, *, freq_mask_width=(0, 30), freq_mask_num=2, time_mask_width=(0, 0.05), time_mask_num=2,
            feat_norm=True, **_):
    m = _get("SpecAugment", input_size=feat.size(-1), time_warp=False, freq_mask_width=list(freq_mask_width),
             freq_mask_num=freq_mask_num, time_mask_width=list(time_mask_width), time_mask_num=time_mask_num,
             feat_norm=feat_norm)
    m.train()
    return m(feat, feat_len)[0]

"""ctypes binding of libsfe.so (the C ABI declared in include/sfe.h).

The library is the product: there is no Python / torch fallback for any entry point.  If the
shared object is missing or fails to load, importing the hot-path modules raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFE_LIB_PATH lets a benchmark load an alternative build of the same ABI (A/B timing on one box)
LIB_PATH = os.environ.get("SFE_LIB_PATH") or os.path.join(_HERE, "csrc", "libsfe.so")

ABI_VERSION = 1

SFE_OK, SFE_ERR_BAD_CONFIG, SFE_ERR_UNSUPPORTED_NFFT, SFE_ERR_BAD_ARG, SFE_ERR_CUDA = range(5)


class SfeConfig(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32), ("center", C.c_int32),
        ("has_preemphasis", C.c_int32), ("preemphasis", C.c_float), ("mag_spec", C.c_int32),
        ("lin_logging", C.c_int32), ("lin_clamp", C.c_float), ("lin_log_div", C.c_float),
        ("n_mels", C.c_int32), ("mel_logging", C.c_int32), ("mel_clamp", C.c_float), ("mel_log_div", C.c_float),
        ("window_host", C.c_void_p), ("mel_weight_host", C.c_void_p),
    ]


class SfeMasks(C.Structure):
    _fields_ = [
        ("freq_pos", C.c_void_p), ("freq_len", C.c_void_p), ("n_freq", C.c_int32),
        ("time_pos", C.c_void_p), ("time_len", C.c_void_p), ("n_time", C.c_int32),
        ("fill", C.c_float), ("fill_dev", C.c_void_p), ("on_host", C.c_int32),
    ]


MASK_INLINE_MAX = 2304


class SfeWarp(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("center", C.c_int32), ("pos", C.c_int32), ("minlen", C.c_int32)]


# every symbol include/sfe.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "sfe_abi_version": (C.c_int, []),
    "sfe_last_error": (C.c_char_p, []),
    "sfe_supported_nfft": (C.c_int, [C.POINTER(C.c_int32), C.c_int]),
    "sfe_plan_create": (C.c_int, [C.POINTER(SfeConfig), C.POINTER(C.c_void_p)]),
    "sfe_plan_destroy": (None, [_vp]),
    "sfe_plan_num_frames": (_i64, [_vp, _i64]),
    "sfe_plan_query": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "sfe_mel_schedule": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32]),
    "sfe_frontend_forward": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       C.POINTER(SfeMasks), _i32, _vp, _vp]),
    "sfe_chain_forward": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _f32, _i32, _i32,
                                    _vp, _vp, C.POINTER(SfeMasks), _vp, _vp, C.POINTER(SfeWarp), _vp]),
    "sfe_running_normalize": (C.c_int, [_vp, _i32, _f32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i64, _i32, _f32,
                                        C.POINTER(SfeMasks), _vp, C.POINTER(SfeWarp), _vp]),
    "sfe_linear2mel_forward": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "sfe_stats_accumulate": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "sfe_stats_finalize": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "sfe_running_update": (C.c_int, [_vp, _i32, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sfe_normalize": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _f32, _i32, _i32, _vp, _vp, _i32, _vp, _vp,
                                _i32, _f32, C.POINTER(SfeWarp), C.POINTER(SfeMasks), _vp]),
    "sfe_prestft_norm": (C.c_int, [_vp, _i64, _vp, _i32, _i64, _i32, _f32, _i32, _vp, _vp, _i64, _vp]),
    "sfe_normalize_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _f32, _i32, _i32, _vp, _vp, _i32,
                                     _i32, _f32, C.POINTER(SfeWarp), C.POINTER(SfeMasks), _vp]),
    "sfe_upload_waveforms": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i64, _vp]),
    "sfe_upload_bytes": (C.c_int64, [_vp, _i32]),
    "sfe_upload_i16_stage_bytes": (C.c_int64, [_i32, _i64]),
    "sfe_upload_waveforms_i16": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i64, _vp, _i64, _vp]),
    "sfe_twosided": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp]),
    "sfe_delta": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp]),
    "sfe_mean_all": (C.c_int, [_vp, _i64, _vp, _vp]),
    "sfe_mel2linear": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _f32, _i64, _f32, _vp, _vp]),
    "sfe_griffinlim_workspace": (_i64, [_vp, _i32, _i64]),
    "sfe_griffinlim": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


class SfeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsfe error {code}: {msg}")
        self.code = code


def load():
    """Load libsfe.so (once) and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library is not built.  Run "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m speechain_b200.csrc.build`). "
                f"There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.sfe_abi_version() != ABI_VERSION:
            raise ImportError(f"libsfe ABI {lib.sfe_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(rc: int):
    if rc != SFE_OK:
        msg = load().sfe_last_error().decode("utf-8", "replace")
        if rc == SFE_ERR_UNSUPPORTED_NFFT or rc == SFE_ERR_BAD_CONFIG:
            raise ValueError(f"libsfe: {msg}")
        raise SfeError(rc, msg)


def supported_nfft():
    buf = (C.c_int32 * 64)()
    n = load().sfe_supported_nfft(buf, 64)
    return [int(buf[i]) for i in range(min(n, 64))]

"""The C-ABI library loads without a GPU and exports every symbol include/sfe.h declares."""
import ctypes
import os
import re

from speechain_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(sfe_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libsfe.so is not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/sfe.h but not exported"


def test_load_types_and_version():
    lib = _lib.load()
    assert lib.sfe_abi_version() == _lib.ABI_VERSION
    nfft = _lib.supported_nfft()
    for n in (400, 512, 1024):  # ASR recipe default, BASELINE C1/C2, TTS recipes
        assert n in nfft


def test_bad_config_is_reported_without_a_gpu():
    # argument validation happens before any CUDA call: error code + message, no exception across the ABI
    cfg = _lib.SfeConfig()
    cfg.n_fft, cfg.win_length, cfg.hop_length = 300, 300, 100
    win = (ctypes.c_float * 300)()
    cfg.window_host = ctypes.cast(win, ctypes.c_void_p)
    h = ctypes.c_void_p()
    rc = _lib.load().sfe_plan_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == _lib.SFE_ERR_UNSUPPORTED_NFFT
    assert b"unsupported n_fft 300" in _lib.load().sfe_last_error()
    cfg.n_fft = 511
    assert _lib.load().sfe_plan_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.SFE_ERR_BAD_CONFIG


def test_new_entry_points_validate_their_arguments_without_a_gpu():
    """sfe_twosided / sfe_running_normalize / sfe_chain_forward refuse inconsistent arguments before any CUDA call."""
    lib = _lib.load()
    one = ctypes.c_void_p(16)  # a non-NULL address that is never dereferenced on these paths
    # two-sided spectrogram: K must be n_fft / 2 + 1, pointers non-NULL, the clamp a normal positive number
    assert lib.sfe_twosided(None, one, 2, 10, 257, 512, 0, 0, 1e-10, 1.0, one, None, None) == _lib.SFE_ERR_BAD_ARG
    assert lib.sfe_twosided(one, one, 2, 10, 256, 512, 0, 0, 1e-10, 1.0, one, None, None) == _lib.SFE_ERR_BAD_ARG
    assert b"bad shape" in lib.sfe_last_error()
    assert lib.sfe_twosided(one, one, 2, 10, 257, 512, 0, 1, 0.0, 1.0, one, None, None) == _lib.SFE_ERR_BAD_ARG
    assert b"clamp" in lib.sfe_last_error()
    # running average + normalise: packed sums, at least one of the two running vectors, a destination; a warp never in place
    args = dict(D=80, weight=0.5, first=0, counter=None, count=2, B=2, F=10, src_has_pad=0, pad=0.0)

    def call(packed, mean, std, feat, out, warp=None):
        return lib.sfe_running_normalize(packed, args["D"], args["weight"], args["first"], mean, std, args["counter"],
                                         args["count"], feat, one, args["B"], args["F"], args["src_has_pad"], args["pad"],
                                         None, None, warp, out)
    assert call(None, one, one, one, one) == _lib.SFE_ERR_BAD_ARG
    assert call(one, None, None, one, one) == _lib.SFE_ERR_BAD_ARG
    assert call(one, one, one, one, None) == _lib.SFE_ERR_BAD_ARG
    assert b"destination" in lib.sfe_last_error()
    w = _lib.SfeWarp(1, 3, 4, 8)
    assert call(one, one, one, one, one, ctypes.byref(w)) == _lib.SFE_ERR_BAD_ARG
    assert b"in place" in lib.sfe_last_error()
    # chain: a NULL plan is refused whatever the mode
    assert lib.sfe_chain_forward(None, one, 100, one, 1, 100, one, one, one, 0, 3, None, None, 1e-10, 1, 1, one, None,
                                 None, None, None, None, None) == _lib.SFE_ERR_BAD_ARG

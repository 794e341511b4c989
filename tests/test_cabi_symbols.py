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

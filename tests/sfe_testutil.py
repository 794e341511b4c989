"""Shared helpers for the test-suite (imported as ``sfe_testutil``)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# North-star tolerance: within 1e-4 relative / 1e-3 absolute (allclose-style) in fp32.
RTOL, ATOL = 1e-4, 1e-3


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def assert_close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert (nan_a == nan_b).all(), f"{what}: NaN pattern differs"
    if a.size == 0:
        return 0.0
    err = np.where(nan_a, 0.0, np.abs(a - b))
    tol = atol + rtol * np.where(nan_b, 0.0, np.abs(b))
    excess = err - tol
    assert (excess <= 0).all(), (
        f"{what}: max abs err {err.max():.3e}, worst excess {excess.max():.3e} at "
        f"{np.unravel_index(np.argmax(excess), a.shape)}")
    return float(err.max())


def assert_gl_close(wav, ref, what=""):
    """Waveforms out of Griffin-Lim: rms error <= 5e-4 of the peak, max error <= 5e-3 of the peak."""
    wav = np.asarray(wav, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert wav.shape == ref.shape, f"{what}: shape {wav.shape} vs {ref.shape}"
    assert np.isfinite(wav).all(), f"{what}: non-finite samples"
    peak = np.abs(ref).max()
    err = np.abs(wav - ref)
    rms = np.sqrt((err ** 2).mean())
    assert err.max() <= 5e-3 * peak and rms <= 5e-4 * peak, (
        f"{what}: max err {err.max():.3e}, rms err {rms:.3e}, peak {peak:.3e}")

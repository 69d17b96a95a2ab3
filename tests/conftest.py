"""Shared test plumbing.  `-m "not gpu"` runs on the CPU build box; `-m gpu` needs a B200."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "weather-radar-processing_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wrp():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def golden_fixtures():
    return np.load(os.path.join(GOLDEN, "fixtures.npz"))


@pytest.fixture(scope="session")
def golden_ref_run():
    return np.load(os.path.join(GOLDEN, "ref_run_sector0.npz"))


# ---- tolerances of SURVEY.md §8d ---------------------------------------------------------
REL_L2 = 1e-4       # per stage: ||x-ref||2 / ||ref||2
ROW_MAX_REL = 1e-4  # per stage: max over rows of max|x-ref| / max|ref|
DB_TOL = 0.01       # |dZdB|, |dZDR| in dB


def rel_l2(x, ref):
    x = np.asarray(x)
    ref = np.asarray(ref)
    return float(np.linalg.norm((x - ref).ravel()) / np.linalg.norm(ref.ravel()))


def line_max_rel(x, ref, axis=-1, floor=1e-3):
    """max |x-ref| / max|ref| per line along `axis`; a line's denominator is floored at
    `floor` x the plane maximum (lines more than 60 dB below the strongest one are judged
    against that floor: fp32 FFT error scales with the strongest line feeding it, finding 4)."""
    x = np.atleast_2d(np.asarray(x))
    ref = np.atleast_2d(np.asarray(ref))
    den = np.abs(ref).max(axis=axis, keepdims=True)
    plane = np.abs(ref).max(axis=(-2, -1), keepdims=True)
    den = np.maximum(den, floor * plane)
    den = np.where(den == 0, 1.0, den)
    return float((np.abs(x - ref) / den).max())


def row_max_rel(x, ref):
    return line_max_rel(x, ref, axis=-1)


def assert_stage_close(x, ref, name="", axis=-1):
    """relL2 <= 1e-4 and line-max-relative <= 1e-4 along the stage's transform axis (rows, or
    columns for the range FFT).  The post-shift DC column is rounding noise in every
    implementation and is covered by the line-max form."""
    assert x.shape == ref.shape, f"{name}: shape {x.shape} vs {ref.shape}"
    l2, rm = rel_l2(x, ref), line_max_rel(x, ref, axis=axis)
    assert l2 <= REL_L2, f"{name}: relL2 {l2:.3e} > {REL_L2}"
    assert rm <= ROW_MAX_REL, f"{name}: line-max-rel {rm:.3e} > {ROW_MAX_REL}"


def assert_products_close(out, zdb_ref, zdr_ref, name=""):
    """out[gates, 2] vs reference ZdB/ZDR: gate 0 must be -inf in both, the rest within 0.01 dB."""
    out = np.asarray(out, dtype=np.float64)
    assert np.isneginf(out[0, 0]) and np.isneginf(zdb_ref[0]), f"{name}: gate 0 must be -inf"
    dz = np.max(np.abs(out[1:, 0] - zdb_ref[1:]))
    dr = np.max(np.abs(out[:, 1] - zdr_ref))
    assert dz <= DB_TOL, f"{name}: max|dZdB| {dz:.3e} dB"
    assert dr <= DB_TOL, f"{name}: max|dZDR| {dr:.3e} dB"

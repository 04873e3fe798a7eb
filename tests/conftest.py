import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    class G:
        kat = np.load(os.path.join(GOLDEN, "kat.npz"))
        fixtures = np.load(os.path.join(GOLDEN, "fixtures.npz"))
        synthetic = np.load(os.path.join(GOLDEN, "synthetic.npz"))
        edge = np.load(os.path.join(GOLDEN, "edge.npz"))
    return G


def rel_err(a, b):
    """max |a-b|/|b| over rows where b is finite (0 when none)."""
    m = np.isfinite(b)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


def assert_parity(got, ref, truth, mode, label="", x_tol=1e-9, o_tol=1e-9):
    """The acceptance rule of SURVEY.md 7/0 and BASELINE.md section 4.

    NaN masks equal exactly.  X-mode: |got-ref|/|ref| <= 1e-9.  O-mode: |got-truth|/|truth|
    <= 1e-9 and |got-ref| <= |ref-truth| + 1e-9 |truth| (inside the reference's own rounding
    ball), because the float64 reference is itself only good to ~1e-5 in O-mode.
    """
    got, ref, truth = (np.asarray(v, dtype=float) for v in (got, ref, truth))
    assert got.shape == ref.shape, label
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "%s: NaN mask differs at %s" % (
        label, np.flatnonzero(np.isnan(got) != np.isnan(ref))[:8])
    m = np.isfinite(ref)
    if not m.any():
        return
    if mode == 'X':
        e = rel_err(got, ref)
        assert e <= x_tol, "%s: X-mode rel err vs reference %.3e" % (label, e)
    else:
        e = rel_err(got, truth)
        assert e <= o_tol, "%s: O-mode rel err vs truth %.3e" % (label, e)
        ball = np.abs(ref[m] - truth[m]) + o_tol * np.abs(truth[m])
        assert np.all(np.abs(got[m] - ref[m]) <= ball), "%s: outside the reference's rounding ball" % label

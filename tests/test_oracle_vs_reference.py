"""CPU, dev container only: the numpy oracle bit-for-bit against the LIVE reference.

Skipped where /root/reference is not mounted (the GPU box); tests/test_oracle_golden.py
covers the same ground through the committed fixtures there.
"""
import warnings

import numpy as np
import pytest

from oracle import vfo_oracle
from oracle.ref_import import load_reference_library, reference_available
from pyrayhf_b200 import synth

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference mount absent")
warnings.simplefilter("ignore")


@pytest.fixture(scope="module")
def ref():
    return load_reference_library()


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n", [1, 3, 64, 200, 1500])
def test_random_synthetic_profiles_bit_exact(ref, mode, n):
    rng = np.random.default_rng(n * 7 + (mode == 'X'))
    lat = rng.uniform(-90, 90, 6)
    lon = rng.uniform(-180, 180, 6)
    alt = synth.default_alt()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    f = np.sort(rng.uniform(0.05, 16.0, 40))
    for p in range(lat.size):
        a = ref.vertical_forward_operator(f, den[p], bmag[p], bpsi[p], alt, mode, n)
        b = vfo_oracle.vertical_forward_operator(f, den[p], bmag[p], bpsi[p], alt, mode, n)
        assert np.array_equal(a, b, equal_nan=True)


def test_stage_functions_bit_exact(ref):
    den, bmag, bpsi, alt = synth.single_day_profile()
    f_hz = synth.default_freq() * 1e6
    assert np.array_equal(ref.smooth_nonuniform_grid(0, 1, 777, 10.), vfo_oracle.stretch_multiplier(777))
    for mode in "OX":
        want = ref.regrid_to_nonuniform_grid(f_hz, den, bmag, bpsi, alt, mode, 300)
        got = vfo_oracle.regrid(f_hz, den, bmag, bpsi, alt, mode, 300)
        for a, b in (("alt", "h"), ("dist", "dh"), ("den", "den"), ("bmag", "bmag"), ("bpsi", "bpsi")):
            assert np.array_equal(want[a], got[b], equal_nan=True), (mode, a)
        assert np.array_equal(want["crit_height"][:, 0], got["h_c"], equal_nan=True)
    X = np.linspace(0.01, 0.99, 50)
    Y = np.full(50, 0.2)
    psi = np.linspace(0, 90, 50)
    for mode in "OX":
        a = ref.find_mu_mup(X, Y, psi, mode)
        b = vfo_oracle.appleton_hartree(X, Y, psi, mode)
        assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1], equal_nan=True)
    assert np.array_equal(ref.find_X(den, 5e6), vfo_oracle.plasma_ratio_x(den, 5e6))
    assert np.array_equal(ref.find_Y(5e6, bmag), vfo_oracle.gyro_ratio_y(5e6, bmag))
    assert ref.constants()[:2] == (vfo_oracle.CP_HZ_PER_SQRT_M3, vfo_oracle.GYRO_HZ_PER_T)


def test_errors_match(ref):
    den, bmag, bpsi, alt = synth.single_day_profile()
    f = np.array([2.0])
    for impl in (ref.vertical_forward_operator, vfo_oracle.vertical_forward_operator):
        with pytest.raises(ValueError):
            impl(f, den, bmag, bpsi, alt, 'Q', 10)
        with pytest.raises(IndexError):
            k = int(np.argmax(den))
            impl(f, den[k:], bmag[k:], bpsi[k:], alt[k:], 'O', 10)
        neg = den.copy()
        neg[0] = -5.0
        with pytest.raises(ValueError):
            impl(f, neg, bmag, bpsi, alt, 'X', 10)


def test_install_rebinds_reference_global(ref):
    """pyrayhf_b200.install() swaps the module global that model_VH resolves at call time (lib:589)."""
    import pyrayhf_b200
    original = ref.vertical_forward_operator
    try:
        pyrayhf_b200.install()
        assert ref.vertical_forward_operator is pyrayhf_b200.vertical_forward_operator
        assert ref.model_VH.__globals__["vertical_forward_operator"] is pyrayhf_b200.vertical_forward_operator
    finally:
        pyrayhf_b200.uninstall()
    assert ref.vertical_forward_operator is original


def test_install_stages_rebinds_every_stage_and_uninstall_restores(ref):
    """install(stages=True) also swaps the standalone stage functions; uninstall() puts every original back."""
    import pyrayhf_b200
    names = ("vertical_forward_operator", "den2freq", "find_X", "find_Y", "smooth_nonuniform_grid",
             "regrid_to_nonuniform_grid", "find_mu_mup", "find_vh")
    originals = {n: getattr(ref, n) for n in names}
    try:
        pyrayhf_b200.install(stages=True)
        for n in names:
            assert getattr(ref, n) is getattr(pyrayhf_b200, n), n
        # the reference's own callers resolve the names through the module globals
        assert ref.find_vh.__module__.startswith("pyrayhf_b200")
        assert ref.vertical_forward_operator.__globals__ is not ref.model_VH.__globals__
        assert ref.model_VH.__globals__["find_X"] is pyrayhf_b200.find_X
    finally:
        pyrayhf_b200.uninstall()
    for n in names:
        assert getattr(ref, n) is originals[n], n
    # the same signatures as the reference (positional names and defaults)
    import inspect
    for n in names + ("constants", "trace_ray_cartesian_snells"):
        want = [(k, v.default) for k, v in inspect.signature(getattr(ref, n)).parameters.items()]
        got = [(k, v.default) for k, v in inspect.signature(getattr(pyrayhf_b200, n)).parameters.items()
               if v.kind is not inspect.Parameter.KEYWORD_ONLY]
        assert got == want, (n, got, want)
    # the spherical tracer's controls are keyword-only in the reference as well (lib:1468-1474)
    want = [(k, v.default, v.kind) for k, v in inspect.signature(ref.trace_ray_spherical_snells).parameters.items()]
    got = [(k, v.default, v.kind) for k, v in inspect.signature(pyrayhf_b200.trace_ray_spherical_snells).parameters.items()]
    assert got == want, (got, want)


def test_residual_tail_matches_reference_residual_VH(ref):
    """oracle.residual_from_model against the reference's residual_VH with model_VH patched to return a given
    curve (the reference's own tests patch the same module globals, tests/test_core.py:345-349)."""
    from unittest.mock import patch

    class P:                                            # minimal stand-in for lmfit.Parameters entries
        def __init__(self, v):
            self.value = v

    params = {"NmF2": P(1e12), "hmF2": P(300.0), "B_bot": P(40.0)}
    F2 = {"Nm": np.array([[1e12]]), "hm": np.array([[300.0]]), "B_bot": np.array([[40.0]])}
    rng = np.random.default_rng(3)
    vh_obs = 200.0 + 100.0 * rng.random(12)
    for frac_nan in (0.0, 0.3, 1.0):
        vh_model = 150.0 + 200.0 * rng.random(12)
        vh_model[rng.random(12) < frac_nan] = np.nan
        with patch("PyRayHF.library.model_VH", return_value=(vh_model.copy(), None)):
            want = ref.residual_VH(params, F2, {}, {}, np.arange(12.0), vh_obs, None, None, None)
        got = vfo_oracle.residual_from_model(vh_obs, vh_model)
        assert np.array_equal(want, got, equal_nan=True)

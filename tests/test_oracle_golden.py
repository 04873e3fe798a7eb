"""CPU: the oracle (numpy restatement + scalar C restatement) against the committed goldens.

The goldens were produced by the live reference (tests/make_golden.py); this pins the
oracle wherever the reference mount is absent (e.g. the GPU box).
"""
import warnings

import numpy as np
import pytest

from conftest import assert_parity, rel_err
from oracle import scalar, vfo_oracle
from pyrayhf_b200 import synth

warnings.simplefilter("ignore")


def test_kat_find_mu_mup(golden):
    k = golden.kat
    mu, mup = vfo_oracle.appleton_hartree(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], 'O')
    # the reference's own pinned literals (tests/test_core.py:137-152), its tolerance
    np.testing.assert_allclose(mu, k["mumup_expected_mu"], rtol=1e-5)
    np.testing.assert_allclose(mup, k["mumup_expected_mup"], rtol=1e-5)
    # and bit-for-bit what the reference computes here
    assert np.array_equal(mu, k["mumup_mu_O"]) and np.array_equal(mup, k["mumup_mup_O"])
    _, mupx = vfo_oracle.appleton_hartree(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], 'X')
    assert np.array_equal(mupx, k["mumup_mup_X"], equal_nan=True)
    # scalar C: literal and long-double truth
    for mode in "OX":
        ref = k["mumup_mup_" + mode]
        assert rel_err(scalar.mup(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], mode, 0), ref) < 1e-12
        assert rel_err(scalar.mup(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], mode, 1), ref) < 1e-11


def test_kat_basic_mask_and_model_vh(golden):
    k = golden.kat
    for mode in "OX":
        vh = vfo_oracle.vertical_forward_operator(k["basic_freq"], k["basic_den"], k["basic_bmag"],
                                                  k["basic_bpsi"], k["basic_alt"], mode, 50)
        assert np.array_equal(vh, k["basic_vh_" + mode], equal_nan=True)
        assert np.isnan(vh[-1]) and np.all(np.isfinite(vh[:-1]))          # tests/test_core.py:235-236
    bm, bp = k["basic_bmag"], k["basic_bpsi"]
    vh = vfo_oracle.vertical_forward_operator(k["model_freq"], k["model_edp"], bm, bp, k["basic_alt"], 'O', 200)
    assert np.array_equal(vh, k["model_vh_O"])
    np.testing.assert_allclose(vh, k["model_expected_vh"], rtol=1e-6)      # tests/test_core.py:275
    fv = vfo_oracle.appleton_hartree(np.array([[0.5, 0.6]]), np.array([[0.1, 0.2]]), np.array([[45.0, 45.0]]), 'O')[1]
    assert abs((np.nansum(fv, axis=1) + 100.0)[0] - k["findvh"][0]) == 0.0


@pytest.mark.parametrize("n", [1, 2, 10, 200, 20000])
def test_multiplier(golden, n):
    m = vfo_oracle.stretch_multiplier(n)
    assert np.array_equal(m, golden.kat["multiplier_%d" % n])
    mc = scalar.multiplier(n)
    assert np.max(np.abs(mc - m)) < 5e-16
    if n > 1:
        assert m[0] == 0.0 and m[-1] == 1.0 and np.all(np.diff(m) > 0)    # tests/test_core.py:181-188


@pytest.mark.parametrize("which", ["Day", "Night"])
@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n", [1, 2, 50, 200, 2000])
def test_fixtures_numpy_oracle_bit_exact(golden, which, mode, n):
    fx = golden.fixtures
    vh = vfo_oracle.vertical_forward_operator(fx["freq_a"], fx[which + "_den"], fx[which + "_bmag"],
                                              fx[which + "_bpsi"], fx[which + "_alt"], mode, n)
    assert np.array_equal(vh, fx["ref_%s_%s_%d_a" % (which, mode, n)], equal_nan=True)


@pytest.mark.parametrize("which", ["Day", "Night"])
@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n,fk", [(200, "a"), (200, "b"), (2000, "a"), (20000, "a"), (20000, "b")])
def test_fixtures_scalar_oracle(golden, which, mode, n, fk):
    fx = golden.fixtures
    tag = "%s_%s_%d_%s" % (which, mode, n, fk)
    args = (fx["freq_" + fk], fx[which + "_den"], fx[which + "_bmag"], fx[which + "_bpsi"], fx[which + "_alt"])
    m = vfo_oracle.stretch_multiplier(n)
    lit, hc = scalar.vertical_forward_operator(*args, mode, n, variant=0, multiplier=m, n_threads=0, return_hc=True)
    tru = scalar.vertical_forward_operator(*args, mode, n, variant=1, multiplier=m, n_threads=0)
    ref = fx["ref_" + tag]
    assert np.array_equal(np.isnan(lit), np.isnan(ref))
    assert np.array_equal(hc, fx["hc_" + tag], equal_nan=True)
    assert np.array_equal(tru, fx["truth_" + tag], equal_nan=True)
    # literal float64 restatement: libm pow/exp differ from numpy's SIMD ones by <= 2 ulp, which is
    # invisible in X-mode and amplified to the reference's own noise level in O-mode
    if mode == 'X':
        assert rel_err(lit, ref) < 1e-11
    else:
        assert rel_err(lit, ref) <= 2.5 * rel_err(ref, tru) + 1e-9
    # libm multiplier instead of numpy's: same to 1e-11 in X-mode
    if mode == 'X':
        lit2 = scalar.vertical_forward_operator(*args, mode, n, variant=0, n_threads=0)
        assert rel_err(lit2, ref) < 1e-11


@pytest.mark.parametrize("mode", ["O", "X"])
def test_synthetic_scalar_oracle(golden, mode):
    sy = golden.synthetic
    den, bmag, bpsi = synth.profiles_at(sy["lat"], sy["lon"], sy["alt"])
    np.testing.assert_allclose([den.sum(), bmag.sum(), bpsi.sum()], sy["input_checksum"], rtol=1e-13)
    m = vfo_oracle.stretch_multiplier(200)
    tru, st = scalar.vertical_forward_operator_batched(sy["freq"], den, bmag, bpsi, sy["alt"], mode, 200,
                                                       variant=1, multiplier=m)
    assert not st.any()
    assert np.array_equal(tru, sy["truth_%s_200" % mode], equal_nan=True)
    lit, _ = scalar.vertical_forward_operator_batched(sy["freq"], den, bmag, bpsi, sy["alt"], mode, 200,
                                                      variant=0, multiplier=m)
    ref = sy["ref_%s_200" % mode]
    assert np.array_equal(np.isnan(lit), np.isnan(ref))
    if mode == 'X':
        assert rel_err(lit, ref) < 1e-10
    # numpy oracle on a handful of profiles, bit-exact
    for p in (0, 17, 250, 251, 255):
        vh = vfo_oracle.vertical_forward_operator(sy["freq"], den[p], bmag[p], bpsi[p], sy["alt"], mode, 200)
        assert np.array_equal(vh, ref[p], equal_nan=True)


EDGE_CASES = ["b_zero", "nan_bmag", "odd_freq", "int_alt", "nonuniform_alt", "valley", "psi_jump",
              "near_crit", "two_level", "peak_at_one"]


@pytest.mark.parametrize("name", EDGE_CASES)
@pytest.mark.parametrize("mode", ["O", "X"])
def test_edge_cases(golden, name, mode):
    e = golden.edge
    args = tuple(e["%s_%s" % (name, k)] for k in ("freq", "den", "bmag", "bpsi", "alt"))
    n = int(e[name + "_n"])
    ref = e["%s_ref_%s" % (name, mode)]
    vh = vfo_oracle.vertical_forward_operator(*args, mode, n)
    assert np.array_equal(vh, ref, equal_nan=True)
    m = vfo_oracle.stretch_multiplier(n)
    lit = scalar.vertical_forward_operator(*args, mode, n, variant=0, multiplier=m)
    tru = scalar.vertical_forward_operator(*args, mode, n, variant=1, multiplier=m)
    assert np.array_equal(np.isnan(lit), np.isnan(ref)), name
    assert np.array_equal(np.isnan(tru), np.isnan(ref)), name
    assert np.array_equal(tru, e["%s_truth_%s" % (name, mode)], equal_nan=True)
    if mode == 'X':
        assert rel_err(lit, ref) < 1e-10


def test_error_behaviour():
    den, bmag, bpsi, alt = synth.single_day_profile()
    f = np.array([2.0, 3.0])
    for impl in (vfo_oracle.vertical_forward_operator, scalar.vertical_forward_operator):
        with pytest.raises(ValueError, match="mode must be 'O' or 'X'"):
            impl(f, den, bmag, bpsi, alt, 'o', 50)
        neg = den.copy()
        neg[3] = -1.0
        with pytest.raises(ValueError, match="Density must be non-negative"):
            impl(f, neg, bmag, bpsi, alt, 'O', 50)
        with pytest.raises(IndexError):
            impl(f, den[::-1].copy()[np.argmax(den[::-1]):], bmag[:den.size - np.argmax(den[::-1])],
                 bpsi[:den.size - np.argmax(den[::-1])], alt[:den.size - np.argmax(den[::-1])], 'O', 50)
    assert vfo_oracle.profile_status(den) == 0
    assert vfo_oracle.profile_status(neg) == 1
    assert vfo_oracle.profile_status(den[np.argmax(den):]) == 2


def test_truth_matches_mpmath_sample(golden):
    """The long-double truth against a 40-digit mpmath evaluation of lib:209-254 on sampled points."""
    import mpmath as mp
    mp.mp.dps = 40
    fx = golden.fixtures
    den, bmag, bpsi, alt = (fx["Night_" + k] for k in ("den", "bmag", "bpsi", "alt"))
    f = np.array([5.0, 9.0, 13.1])
    _, g = vfo_oracle.vertical_forward_operator(f, den, bmag, bpsi, alt, 'O', 200, stages=True)

    def mp_mup(X, Y, psi, sgn):
        X, Y, psi = mp.mpf(float(X)), mp.mpf(float(Y)), mp.mpf(float(psi))
        r = psi * mp.pi / 180
        s, c = mp.sin(r), mp.cos(r)
        YT, YL, Xm1 = Y * s, Y * c, 1 - X
        beta = mp.sqrt(mp.mpf(1) / 4 * YT ** 4 + YL ** 2 * Xm1 ** 2)
        D = Xm1 - YT ** 2 / 2 + sgn * beta
        mu = mp.sqrt(1 - X * Xm1 / D)
        dbdx = -YL ** 2 * Xm1 / beta
        dDdX = -1 + sgn * dbdx
        dady = YT ** 3 * s + 2 * YL * Xm1 ** 2 * c
        dDdY = -YT * s + sgn * dady / (2 * beta)
        dmudY = X * Xm1 * dDdY / (2 * mu * D ** 2)
        dmudX = (2 * X - 1 + X * Xm1 / D * dDdX) / (2 * mu * D)
        return mu - (2 * X * dmudX + Y * dmudY)

    for mode, sgn in (("O", 1), ("X", -1)):
        idx = [0, 57, 150, 198, 199]
        for r in range(3):
            X, Y, P = g["X"][r, idx], g["Y"][r, idx], g["bpsi"][r, idx]
            if mode == 'X':
                X = X * 0.8        # stay below the X-mode cutoff X = 1 - Y
            want = np.array([float(mp_mup(x, y, p, sgn)) for x, y, p in zip(X, Y, P)])
            got = scalar.mup(X, Y, P, mode, variant=1)
            # the last grid point sits 1e-6 km below X = 1: (2X - 1 + q dD/dX) cancels to O(1 - X) ~ 1e-9,
            # so even 64-bit-mantissa arithmetic keeps only ~1e-11 there (its weight in vh is ~1e-5)
            assert rel_err(got[:-1], want[:-1]) < 1e-13, (mode, r)
            assert rel_err(got[-1:], want[-1:]) < 1e-10, (mode, r)

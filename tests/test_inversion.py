"""Inversion objective: ``minimize_parameters``' brute search (library.py:672-825) on the device.

Goldens (tests/golden/inversion.npz) come from the LIVE reference: its own ``minimize_parameters`` + ``residual_VH``
running the real ``scipy.optimize.brute`` through tests/lmfit_shim.py, with ``model_VH`` patched by a Chapman profile
builder the way tests/test_core.py:345-349 patch it (tests/make_golden_inversion.py).
"""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inversion.npz"))


def test_host_side_pieces_match_the_reference():
    """No GPU: observation clean-up (library.py:732-736), NmF2 from the highest frequency (757-778), brute grid."""
    from pyrayhf_b200 import inversion
    for mode in ("O", "X"):
        f_in, vh_obs = inversion.sort_observations(GOLD[mode + "_f_obs"], GOLD[mode + "_vh_obs"])
        assert np.all(np.diff(f_in) > 0) and np.all(np.isfinite(vh_obs)) and f_in.size == GOLD[mode + "_f_obs"].size - 1
        nm = inversion.nmf2_from_max_frequency(f_in[-1], GOLD["alt"], GOLD["bmag"], 310.0, mode)
        assert nm == float(GOLD[mode + "_nmf2"])             # same operations in the same order: bit-equal
        grid = GOLD[mode + "_grid"]                           # [2, 31, 5] from scipy.optimize.brute
        assert np.array_equal(inversion.brute_grid(310.0 - 31.0, 310.0 + 31.0, 2.0), grid[0][:, 0])
        assert np.array_equal(inversion.brute_grid(50.0 - 5.0, 50.0 + 5.0, 2.0), grid[1][0, :])
    assert inversion.freq2den(8.97866275e6) == (8.97866275e6 / 8.97866275) ** 2
    with pytest.raises(ValueError, match="B_bot is not provided"):
        inversion.minimize_parameters({"Nm": 1, "hm": 1}, {}, {}, np.ones(2), np.ones(2), GOLD["alt"], GOLD["bmag"],
                                      GOLD["bpsi"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["X", "O"])
def test_minimize_parameters_brute_matches_the_reference(mode):
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import inversion
    alt, bmag, bpsi = GOLD["alt"], GOLD["bmag"], GOLD["bpsi"]
    builder = inversion.chapman_profile_builder(float(GOLD["e_fo_mhz"]))
    F2 = {"Nm": np.array([[1e12]]), "hm": np.array([[310.0]]), "B_bot": np.array([[50.0]])}
    F1 = {"P": np.array([[0.5]])}
    E = {"hm": np.array([[110.0]])}
    f_obs, vh_obs = GOLD[mode + "_f_obs"], GOLD[mode + "_vh_obs"]
    vh_res, edp, f2_fit = prhf.minimize_parameters(F2, F1, E, f_obs, vh_obs, alt, bmag, bpsi, method='brute',
                                                   percent_sigma=10., step=2., mode=mode, n_points=200,
                                                   profile_builder=builder)
    assert float(np.squeeze(f2_fit["Nm"])) == float(GOLD[mode + "_nmf2"])
    assert float(np.squeeze(f2_fit["hm"])) == float(GOLD[mode + "_hmf2"])
    assert float(np.squeeze(f2_fit["B_bot"])) == float(GOLD[mode + "_b_bot"])
    assert F2["hm"][0, 0] == 310.0                            # inputs are not mutated
    np.testing.assert_allclose(edp, GOLD[mode + "_edp_result"], rtol=1e-12)
    ref = GOLD[mode + "_vh_result"]
    assert vh_res.shape == f_obs.shape and np.array_equal(np.isnan(vh_res), np.isnan(ref))
    m = np.isfinite(ref)
    # X-mode: 1e-9 against the reference; O-mode: the float64 reference carries 1e-6..1e-4 of its own rounding noise
    np.testing.assert_allclose(vh_res[m], ref[m], rtol=1e-9 if mode == "X" else 2e-4)
    # the whole objective surface, node by node (scipy.optimize.brute's Jout, C order: hmF2 outer, B_bot inner)
    f_in, obs = inversion.sort_observations(f_obs, vh_obs)
    grid = GOLD[mode + "_grid"]
    den = builder(float(GOLD[mode + "_nmf2"]), grid[0].ravel(), grid[1].ravel(), alt)
    best, chi2, vh_all = prhf.brute_force_fit(f_in, obs, den, bmag, bpsi, alt, mode, 200)
    jout = GOLD[mode + "_jout"].ravel()
    assert best == int(np.argmin(jout))
    np.testing.assert_allclose(chi2.cpu().numpy(), jout, rtol=1e-8 if mode == "X" else 1e-3)
    b2, c2 = prhf.brute_force_fit(f_in, obs, den, bmag, bpsi, alt, mode, 200, return_arrays=False)
    assert b2 == best and c2 == float(chi2[best])


@pytest.mark.gpu
def test_brute_force_fit_skips_failed_and_dead_candidates():
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    alt = synth.default_alt()
    freq = np.arange(2.0, 9.0, 0.25)
    fof2 = np.linspace(8.0, 12.0, 9)
    hmf2 = np.linspace(280.0, 340.0, 7)
    ff, hh = np.meshgrid(fof2, hmf2, indexing="ij")
    den, bmag, bpsi = synth.profiles_from_parameters(ff.ravel(), hh.ravel(), np.full(ff.size, 50.0),
                                                     np.full(ff.size, 3.0), np.full(ff.size, 20.0), alt)
    truth = 31
    obs = prhf.vertical_forward_operator(freq, den[truth], bmag[truth], bpsi[truth], alt, 'O', 200)
    den = den.copy()
    den[0, 5] = -1.0                                          # the reference raises for this candidate: never selected
    den[1] = 0.0                                              # no reflection at any frequency, peak at index 0
    best, chi2, vh = prhf.brute_force_fit(freq, obs, den, bmag[0], bpsi[0], alt, 'O', 200)
    assert best == truth and chi2[truth] == 0.0 and np.isnan(chi2[0]) and np.isnan(chi2[1])
    assert vh.shape == (ff.size, freq.size)
    bad = np.zeros((3, alt.size))
    assert prhf.brute_force_fit(freq, obs, bad, bmag[0], bpsi[0], alt, 'O', 200, return_arrays=False)[0] == -1

"""GPU: the CUDA path (through the C ABI) against the oracle and the committed goldens.

Acceptance (conftest.assert_parity): NaN masks identical; X-mode <= 1e-9 relative against
the reference; O-mode <= 1e-9 against the long-double truth and inside the reference's own
rounding ball.
"""
import warnings

import numpy as np
import pytest

from conftest import assert_parity, rel_err
from oracle import scalar, vfo_oracle
from pyrayhf_b200 import synth

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")


@pytest.fixture(scope="module")
def vfo():
    import torch
    assert torch.cuda.is_available()
    import pyrayhf_b200
    return pyrayhf_b200


def test_kat(vfo, golden):
    k = golden.kat
    for mode in "OX":
        got = vfo.vertical_forward_operator(k["basic_freq"], k["basic_den"], k["basic_bmag"], k["basic_bpsi"],
                                            k["basic_alt"], mode, 50)
        assert isinstance(got, np.ndarray) and got.dtype == np.float64 and got.shape == (3,)
        ref = k["basic_vh_" + mode]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.isnan(got[-1]) and np.all(np.isfinite(got[:-1]))
        if mode == 'X':
            assert rel_err(got, ref) < 1e-9
    got = vfo.vertical_forward_operator(k["model_freq"], k["model_edp"], k["basic_bmag"], k["basic_bpsi"],
                                        k["basic_alt"], 'O', 200)
    # tests/test_core.py:275 pins these O-mode values to rtol 1e-6, but the float64 reference that produced
    # them is itself 3.4e-6 away from an exact evaluation of its own formulas (cancellation in lib:229);
    # the acceptance rule is therefore the rounding-ball one, and the pinned literals hold to 5e-6.
    assert_parity(got, k["model_vh_O"], k["model_truth_O"], 'O', "model_VH KAT")
    np.testing.assert_allclose(got, k["model_expected_vh"], rtol=5e-6)
    np.testing.assert_allclose(k["model_vh_O"], k["model_expected_vh"], rtol=1e-6)


def test_kat_mu_mup(vfo, golden):
    k = golden.kat
    for literal in (False, True):
        mu, mup = vfo.find_mu_mup(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], 'O', literal=literal)
        np.testing.assert_allclose(mu, k["mumup_expected_mu"], rtol=1e-5)   # tests/test_core.py:149-152
        np.testing.assert_allclose(mup, k["mumup_expected_mup"], rtol=1e-5)
        assert rel_err(mup, k["mumup_mup_O"]) < 1e-11
        mux, mupx = vfo.find_mu_mup(k["mumup_X"], k["mumup_Y"], k["mumup_psi"], 'X', literal=literal)
        assert np.array_equal(np.isnan(mupx), np.isnan(k["mumup_mup_X"]))
        assert rel_err(mupx, k["mumup_mup_X"]) < 1e-11


@pytest.mark.parametrize("which", ["Day", "Night"])
@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n,fk", [(1, "a"), (2, "a"), (50, "a"), (200, "a"), (200, "b"), (2000, "a"),
                                  (20000, "a"), (20000, "b")])
def test_tutorial_fixtures(vfo, golden, which, mode, n, fk):
    fx = golden.fixtures
    tag = "%s_%s_%d_%s" % (which, mode, n, fk)
    got = vfo.vertical_forward_operator(fx["freq_" + fk], fx[which + "_den"], fx[which + "_bmag"],
                                        fx[which + "_bpsi"], fx[which + "_alt"], mode, n)
    assert_parity(got, fx["ref_" + tag], fx["truth_" + tag], mode, tag)


@pytest.mark.parametrize("mode", ["O", "X"])
def test_literal_flag_tracks_reference(vfo, golden, mode):
    """The literal path reproduces the reference's operation order: X-mode to 1e-11, O-mode to
    within a small multiple of the reference's own cancellation noise."""
    fx = golden.fixtures
    for which in ("Day", "Night"):
        tag = "%s_%s_200_a" % (which, mode)
        got = vfo.vertical_forward_operator(fx["freq_a"], fx[which + "_den"], fx[which + "_bmag"],
                                            fx[which + "_bpsi"], fx[which + "_alt"], mode, 200, literal=True)
        ref, tru = fx["ref_" + tag], fx["truth_" + tag]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        if mode == 'X':
            assert rel_err(got, ref) < 1e-11
        else:
            assert rel_err(got, ref) <= 2.5 * rel_err(ref, tru) + 1e-9


@pytest.mark.parametrize("mode", ["O", "X"])
def test_synthetic_batch_numpy(vfo, golden, mode):
    sy = golden.synthetic
    den, bmag, bpsi = synth.profiles_at(sy["lat"], sy["lon"], sy["alt"])
    got, st = vfo.vertical_forward_operator_batched(sy["freq"], den, bmag, bpsi, sy["alt"], mode, 200,
                                                    return_status=True)
    assert not st.any()
    assert_parity(got, sy["ref_%s_200" % mode], sy["truth_%s_200" % mode], mode, "synthetic 200")
    sub = sy["sub_20000"]
    got = vfo.vertical_forward_operator_batched(sy["freq"], den[sub], bmag[sub], bpsi[sub], sy["alt"], mode, 20000)
    assert_parity(got, sy["ref_%s_20000" % mode], sy["truth_%s_20000" % mode], mode, "synthetic 20000")


@pytest.mark.parametrize("mode", ["O", "X"])
def test_synthetic_batch_torch_device(vfo, golden, mode):
    import torch
    sy = golden.synthetic
    den, bmag, bpsi = synth.profiles_at(sy["lat"], sy["lon"], sy["alt"])
    dev = torch.device("cuda:0")
    P = den.shape[0]
    # per-profile freq and alt rows exercise the strided variants
    freq2 = np.tile(sy["freq"], (P, 1))
    alt2 = np.tile(sy["alt"], (P, 1))
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq2, den, bmag, bpsi, alt2)]
    vh, st = vfo.vertical_forward_operator_batched(*t, mode, 200, return_status=True)
    assert vh.is_cuda and vh.dtype == torch.float64 and tuple(vh.shape) == (P, sy["freq"].size)
    assert int(st.abs().sum().item()) == 0
    assert_parity(vh.cpu().numpy(), sy["ref_%s_200" % mode], sy["truth_%s_200" % mode], mode, "torch batch")


EDGE_CASES = ["b_zero", "nan_bmag", "odd_freq", "int_alt", "nonuniform_alt", "valley", "psi_jump",
              "near_crit", "near_crit_hi", "two_level", "peak_at_one"]


@pytest.mark.parametrize("name", EDGE_CASES)
@pytest.mark.parametrize("mode", ["O", "X"])
def test_edge_cases(vfo, golden, name, mode):
    e = golden.edge
    args = tuple(e["%s_%s" % (name, k)] for k in ("freq", "den", "bmag", "bpsi", "alt"))
    n = int(e[name + "_n"])
    got = vfo.vertical_forward_operator(*args, mode, n)
    assert_parity(got, e["%s_ref_%s" % (name, mode)], e["%s_truth_%s" % (name, mode)], mode, name)


def test_error_behaviour(vfo):
    den, bmag, bpsi, alt = synth.single_day_profile()
    f = np.array([2.0, 3.0])
    with pytest.raises(ValueError, match="mode must be 'O' or 'X'"):
        vfo.vertical_forward_operator(f, den, bmag, bpsi, alt, 'o', 50)
    neg = den.copy()
    neg[3] = -1.0
    with pytest.raises(ValueError, match="Density must be non-negative"):
        vfo.vertical_forward_operator(f, neg, bmag, bpsi, alt, 'O', 50)
    k = int(np.argmax(den))
    with pytest.raises(IndexError):
        vfo.vertical_forward_operator(f, den[k:], bmag[k:], bpsi[k:], alt[k:], 'O', 50)
    # batched: failed profiles are reported per profile and come back as NaN rows
    lat, lon = synth.grid_subset(5)
    d2, b2, p2 = synth.profiles_at(lat, lon, alt)
    d2[1, 2] = -3.0
    d2[3] = d2[3, ::-1]
    vh, st = vfo.vertical_forward_operator_batched(f, d2, b2, p2, alt, 'X', 100, errors='nan', return_status=True)
    assert list(st) == [0, 1, 0, 2 if np.argmax(d2[3]) == 0 else 0, 0]
    assert np.all(np.isnan(vh[1]))
    ref = vfo_oracle.vertical_forward_operator(f, d2[0], b2[0], p2[0], alt, 'X', 100)
    assert np.array_equal(np.isnan(vh[0]), np.isnan(ref)) and rel_err(vh[0], ref) < 1e-9
    with pytest.raises(ValueError):
        vfo.vertical_forward_operator_batched(f, d2, b2, p2, alt, 'X', 100)
    # float32 / int inputs are up-cast like the reference does
    got = vfo.vertical_forward_operator(f.astype(np.float32), den.astype(np.float32), bmag, bpsi, alt.astype(np.int64), 'X', 100)
    ref = vfo_oracle.vertical_forward_operator(f.astype(np.float32).astype(float), den.astype(np.float32).astype(float),
                                               bmag, bpsi, alt.astype(float), 'X', 100)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and rel_err(got, ref) < 1e-9
    # 0-d frequency -> shape (1,)
    got = vfo.vertical_forward_operator(np.array(2.0), den, bmag, bpsi, alt, 'X', 100)
    assert got.shape == (1,)
    # freq of shape (1, F) broadcasts in the reference and returns (F,); other 2-D / 3-D shapes fail there too
    got2 = vfo.vertical_forward_operator(f[None, :], den, bmag, bpsi, alt, 'X', 100)
    assert got2.shape == f.shape and np.array_equal(got2, vfo.vertical_forward_operator(f, den, bmag, bpsi, alt, 'X', 100),
                                                    equal_nan=True)
    for bad in (f[:, None], np.tile(f, (2, 1)), f[None, None, :]):
        with pytest.raises(ValueError):
            vfo.vertical_forward_operator(bad, den, bmag, bpsi, alt, 'X', 100)
    with pytest.raises(TypeError):
        vfo.vertical_forward_operator(list(f), den, bmag, bpsi, alt, 'X', 100)
    with pytest.raises(AttributeError):
        vfo.vertical_forward_operator(2.0, den, bmag, bpsi, alt, 'X', 100)
    assert vfo.vertical_forward_operator(np.float64(2.0), den, bmag, bpsi, alt, 'X', 100).shape == (1,)


@pytest.mark.parametrize("seg_len", [256, 1024, 4096, 100000])
def test_tiling_invariance(vfo, golden, seg_len, monkeypatch):
    """The result must not depend on how rows are split into tiles (fresh ctx per setting)."""
    from pyrayhf_b200 import _cabi
    monkeypatch.setenv("PRHF_SEG_LEN", str(seg_len))
    monkeypatch.setattr(_cabi, "_contexts", {})
    fx = golden.fixtures
    got = vfo.vertical_forward_operator(fx["freq_a"], fx["Day_den"], fx["Day_bmag"], fx["Day_bpsi"], fx["Day_alt"],
                                        'X', 20000)
    assert_parity(got, fx["ref_Day_X_20000_a"], fx["truth_Day_X_20000_a"], 'X', "seg_len %d" % seg_len)
    # repeated call on the same ctx exercises the self-resetting counters
    got2 = vfo.vertical_forward_operator(fx["freq_a"], fx["Day_den"], fx["Day_bmag"], fx["Day_bpsi"], fx["Day_alt"],
                                         'X', 20000)
    assert np.array_equal(got, got2, equal_nan=True)


def test_full_size_properties(vfo):
    """BASELINE config sizes, size-independent properties instead of an oracle run:
    determinism, batch == single, permutation equivariance over profiles and frequencies."""
    import torch
    lat, lon = synth.grid_subset(64, seed=3)
    alt = synth.default_alt()
    freq = synth.default_freq()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    a = vfo.vertical_forward_operator_batched(*t, 'X', 20000).cpu().numpy()
    b = vfo.vertical_forward_operator_batched(*t, 'X', 20000).cpu().numpy()
    assert np.array_equal(a, b, equal_nan=True)
    perm = np.random.default_rng(0).permutation(64)
    tp = [t[0]] + [v[torch.from_numpy(perm).to(dev)] for v in t[1:4]] + [t[4]]
    c = vfo.vertical_forward_operator_batched(*tp, 'X', 20000).cpu().numpy()
    assert np.array_equal(c, a[perm], equal_nan=True)
    fperm = np.random.default_rng(1).permutation(freq.size)
    tf = [t[0][torch.from_numpy(fperm).to(dev)]] + t[1:]
    d = vfo.vertical_forward_operator_batched(*tf, 'X', 20000).cpu().numpy()
    assert np.array_equal(d, a[:, fperm], equal_nan=True)
    single = vfo.vertical_forward_operator(freq, den[5], bmag[5], bpsi[5], alt, 'X', 20000)
    assert np.array_equal(np.isnan(single), np.isnan(a[5])) and rel_err(single, a[5]) < 1e-13
    # virtual height is >= true height of reflection >= bottom of the profile, and increases
    # with frequency inside each layer: check the weaker invariant vh >= alt[0]
    assert np.nanmin(a) >= alt[0]
    # oracle spot check at full n_points on two profiles
    for p in (0, 63):
        ref = vfo_oracle.vertical_forward_operator(freq, den[p], bmag[p], bpsi[p], alt, 'X', 20000)
        assert np.array_equal(np.isnan(a[p]), np.isnan(ref)) and rel_err(a[p], ref) < 1e-9


def test_fast_math_accuracy(vfo):
    """rcp_fast / rsqrt_fast (MUFU seed + refinement) against IEEE division and sqrt."""
    from pyrayhf_b200 import _cabi
    errs = _cabi.context(0).selftest_math()
    print("fast-math max rel err [rcp, rsqrt, rcp seed, rsqrt seed, rcp cubic, rsqrt cubic]:", errs)
    e_rcp, e_rsqrt = errs[0], errs[1]
    assert 0.0 <= e_rcp < 4.5e-16, e_rcp
    assert 0.0 <= e_rsqrt < 4.5e-16, e_rsqrt


@pytest.mark.parametrize("mode", ["O", "X"])
def test_critical_frequency_ulp_sweep(vfo, golden, mode):
    """Frequencies within a few ulp of making X (or X+Y) == 1 exactly at a profile level: the row-setup
    kernel's screened scan must take every validity / bracket decision exactly as the literal one."""
    fx = golden.fixtures
    den, bmag, bpsi, alt = (fx["Day_" + k] for k in ("den", "bmag", "bpsi", "alt"))
    nt = int(np.argmax(den))
    freqs = []
    for k in (5, 40, 120, nt - 2, nt - 1):
        fp = np.sqrt(den[k]) * 8.97866275                 # plasma frequency at level k [Hz]
        if mode == 'X':
            fh = 2.799249247e10 * bmag[k]
            f0 = 0.5 * (fh + np.sqrt(fh * fh + 4 * fp * fp))   # X + Y = 1
        else:
            f0 = fp
        f0 /= 1e6
        for j in range(-6, 7):
            f = f0
            for _ in range(abs(j)):
                f = np.nextafter(f, np.inf if j > 0 else -np.inf)
            freqs.append(f)
    freqs = np.array(freqs)
    got = vfo.vertical_forward_operator(freqs, den, bmag, bpsi, alt, mode, 300)
    ref = vfo_oracle.vertical_forward_operator(freqs, den, bmag, bpsi, alt, mode, 300)
    tru = scalar.vertical_forward_operator(freqs, den, bmag, bpsi, alt, mode, 300, variant=1,
                                           multiplier=vfo_oracle.stretch_multiplier(300))
    assert np.isnan(ref).any() and np.isfinite(ref).any()
    # rows whose reflection level is a flat spot of the profile are ill-conditioned in h_c itself;
    # the mask must still agree exactly and the values to 1e-9 (X) / rounding ball (O)
    assert_parity(got, ref, tru, mode, "ulp sweep")


@pytest.mark.parametrize("mode", ["O", "X"])
def test_planned_mode_matches_direct_mode(vfo, golden, mode, monkeypatch):
    """Small batches use K1's tile planner (adaptive segments, compact live-tile list); the result must
    not depend on it.  PRHF_PLANNED_MAX_ROWS=0 forces the direct mode."""
    from pyrayhf_b200 import _cabi
    fx = golden.fixtures
    args = (fx["freq_a"], fx["Night_den"], fx["Night_bmag"], fx["Night_bpsi"], fx["Night_alt"], mode, 20000)
    planned = vfo.vertical_forward_operator(*args)
    planned2 = vfo.vertical_forward_operator(*args)
    assert np.array_equal(planned, planned2, equal_nan=True)          # self-resetting counters, determinism
    monkeypatch.setenv("PRHF_PLANNED_MAX_ROWS", "0")
    monkeypatch.setattr(_cabi, "_contexts", {})
    direct = vfo.vertical_forward_operator(*args)
    assert np.array_equal(np.isnan(planned), np.isnan(direct))
    assert rel_err(planned, direct) < 1e-13
    assert_parity(planned, fx["ref_Night_%s_20000_a" % mode], fx["truth_Night_%s_20000_a" % mode], mode, "planned")
    # a handful of profiles (still planned mode), incl. one failed profile and an all-dead one
    monkeypatch.delenv("PRHF_PLANNED_MAX_ROWS")
    monkeypatch.setattr(_cabi, "_contexts", {})
    lat, lon = synth.grid_subset(6, seed=5)
    alt = synth.default_alt()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    den[2, 3] = -1.0
    den[4] *= 1e-6                                    # foF2 far below every sounding frequency
    f = synth.default_freq()
    got, st = vfo.vertical_forward_operator_batched(f, den, bmag, bpsi, alt, mode, 5000, errors='nan',
                                                    return_status=True)
    assert list(st) == [0, 0, 1, 0, 0, 0]
    for p in (0, 1, 3, 4, 5):
        ref = vfo_oracle.vertical_forward_operator(f, den[p], bmag[p], bpsi[p], alt, mode, 5000)
        assert np.array_equal(np.isnan(got[p]), np.isnan(ref)), p
        if mode == 'X':
            assert rel_err(got[p], ref) < 1e-9
    assert np.all(np.isnan(got[2]))


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n", [200, 1500])
def test_row_per_warp_kernel_matches_tile_kernel(vfo, golden, mode, n, monkeypatch):
    """Small n_points go through the row-per-warp kernel; PRHF_NO_ROWWARP=1 forces the tile kernel."""
    from pyrayhf_b200 import _cabi
    sy = golden.synthetic
    den, bmag, bpsi = synth.profiles_at(sy["lat"], sy["lon"], sy["alt"])
    a = vfo.vertical_forward_operator_batched(sy["freq"], den, bmag, bpsi, sy["alt"], mode, n)
    monkeypatch.setenv("PRHF_NO_ROWWARP", "1")
    monkeypatch.setattr(_cabi, "_contexts", {})
    b = vfo.vertical_forward_operator_batched(sy["freq"], den, bmag, bpsi, sy["alt"], mode, n)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    # the two kernels scale density / field at different points of the arithmetic; O-mode amplifies the
    # resulting 1-ulp differences in X near reflection, X-mode does not
    assert rel_err(a, b) < (2e-10 if mode == 'X' else 5e-10)
    if n == 200:
        assert_parity(a, sy["ref_%s_200" % mode], sy["truth_%s_200" % mode], mode, "row-per-warp")


@pytest.mark.parametrize("mode", ["O", "X"])
def test_row_setup_lane_mode_matches_warp_mode(vfo, golden, mode, monkeypatch):
    """Large batches run the row setup with one thread per frequency; the reflection heights (hence every
    result bit) must equal the one-warp-per-frequency scan.  PRHF_NO_K1_LANES=1 forces the latter."""
    from pyrayhf_b200 import _cabi
    sy = golden.synthetic
    den, bmag, bpsi = synth.profiles_at(sy["lat"], sy["lon"], sy["alt"])
    # frequencies within a few ulp of exact crossings of the first profile, to hit the fallback scans too
    k = 60
    fp = np.sqrt(den[0, k]) * 8.97866275
    fh = 2.799249247e10 * bmag[0, k]
    f0 = (0.5 * (fh + np.sqrt(fh * fh + 4 * fp * fp)) if mode == 'X' else fp) / 1e6
    extra = [f0]
    for _ in range(3):
        extra.append(np.nextafter(extra[-1], np.inf))
    freq = np.concatenate([sy["freq"], extra, [np.nextafter(f0, 0.0)]])
    a = vfo.vertical_forward_operator_batched(freq, den, bmag, bpsi, sy["alt"], mode, 200)
    monkeypatch.setenv("PRHF_NO_K1_LANES", "1")
    monkeypatch.setattr(_cabi, "_contexts", {})
    b = vfo.vertical_forward_operator_batched(freq, den, bmag, bpsi, sy["alt"], mode, 200)
    assert np.array_equal(a, b, equal_nan=True)
    assert_parity(a[:, :sy["freq"].size], sy["ref_%s_200" % mode], sy["truth_%s_200" % mode], mode, "lane-mode K1")
    ref = vfo_oracle.vertical_forward_operator(freq, den[0], bmag[0], bpsi[0], sy["alt"], mode, 200)
    assert np.array_equal(np.isnan(a[0]), np.isnan(ref))


@pytest.mark.parametrize("mode", ["O", "X"])
def test_config5_frequency_sweep_full_size(vfo, golden, mode):
    """BASELINE configs[4]: 1740 frequencies (0.01 MHz step) at n_points = 50000 on the tutorial Day profile.
    The numpy oracle would need ~14 GB of temporaries here, so the checker is the scalar C oracle:
    its literal variant is within 1e-11 of the reference in X-mode (test_oracle_golden) and its long-double
    variant is the O-mode truth.  Rows within 0.05 MHz of the truncated-peak cutoff are reported separately."""
    fx = golden.fixtures
    den, bmag, bpsi, alt = (fx["Day_" + k] for k in ("den", "bmag", "bpsi", "alt"))
    freq = np.arange(0.01, 17.41, 0.01)
    n = 50000
    got = vfo.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n)
    m = vfo_oracle.stretch_multiplier(n)
    want = scalar.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n, variant=(0 if mode == 'X' else 1),
                                            multiplier=m, n_threads=0)
    bad = np.flatnonzero(np.isnan(got) != np.isnan(want))
    assert bad.size == 0, "mask mismatch at rows %s: got %s want %s" % (bad[:8], got[bad[:8]], want[bad[:8]])
    assert rel_err(got, want) < 1e-9
    k = int(np.argmax(den))
    fcut = np.sqrt(den[k - 1]) * 8.97866275 / 1e6
    near = np.abs(freq - fcut) < 0.05
    assert near.any()
    print("config5 %s: %d finite rows, max rel err %.2e (near-cutoff rows: %.2e)" % (
        mode, int(np.isfinite(want).sum()), rel_err(got, want), rel_err(got[near], want[near])))


def test_residual_objective(vfo):
    """prhf_residual_f64 (tail of residual_VH, lib:660-668) against its numpy restatement, and the batched
    brute-force scoring built on it."""
    import torch
    rng = np.random.default_rng(11)
    n_prof, n_freq = 300, 37
    vh_obs = 200.0 + 100.0 * rng.random(n_freq)
    vh = 150.0 + 200.0 * rng.random((n_prof, n_freq))
    vh[rng.random((n_prof, n_freq)) < 0.25] = np.nan
    vh[7] = np.nan                                            # all-NaN row: nanmean is NaN, residual all NaN
    vh[8] = 50.0                                              # finite, mean below 100 (fill unused)
    vh[9, ::2] = np.nan
    vh[9, 1::2] = 20.0                                        # mean 20 -> fill 100
    want = vfo_oracle.residual_from_model(vh_obs, vh)
    res, chi2 = vfo.residual_VH_batched(vh_obs, vh)
    assert np.array_equal(np.isnan(res), np.isnan(want))
    # the NaN fill is a mean (different summation order on the GPU: last-ulp differences), and residuals are
    # differences of ~300 km numbers: compare absolutely
    m = np.isfinite(want)
    assert np.max(np.abs(res[m] - want[m])) < 1e-10
    np.testing.assert_allclose(chi2[np.isfinite(chi2)], np.sum(want ** 2, axis=1)[np.isfinite(chi2)], rtol=1e-10)
    assert np.isnan(chi2[7]) and np.all(np.isnan(res[7]))
    # torch in / torch out
    dev = torch.device("cuda:0")
    r2, c2 = vfo.residual_VH_batched(torch.from_numpy(vh_obs).to(dev), torch.from_numpy(vh).to(dev))
    assert np.array_equal(r2.cpu().numpy(), res, equal_nan=True) and np.array_equal(c2.cpu().numpy(), chi2, equal_nan=True)


def test_config3_global_grid_full_size(vfo):
    """BASELINE configs[2]: the 181 x 361 one-degree grid (65,341 profiles), O and X mode, 174 frequencies,
    n_points = 200, as CUDA tensors.  Checked through size-independent properties (no status errors, determinism,
    each sampled row equals the single-profile call bit for bit) and against the oracle on sampled profiles."""
    import torch
    lat, lon = synth.global_grid_points()
    alt = synth.default_alt()
    freq = synth.default_freq()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    assert den.shape == (65341, 620)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    rng = np.random.default_rng(5)
    sample = np.concatenate([[0, 180, 65340, 32670], rng.integers(0, 65341, 12)])
    for mode in "OX":
        vh, st = vfo.vertical_forward_operator_batched(*t, mode, 200, return_status=True)
        assert int(st.abs().sum().item()) == 0
        a = vh.cpu().numpy()
        b = vfo.vertical_forward_operator_batched(*t, mode, 200).cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True)
        assert np.nanmin(a) >= alt[0] and np.isfinite(a).mean() > 0.2
        m = vfo_oracle.stretch_multiplier(200)
        for p in sample:
            ref = vfo_oracle.vertical_forward_operator(freq, den[p], bmag[p], bpsi[p], alt, mode, 200)
            tru = scalar.vertical_forward_operator(freq, den[p], bmag[p], bpsi[p], alt, mode, 200, variant=1,
                                                   multiplier=m)
            assert_parity(a[p], ref, tru, mode, "config3 profile %d" % p)


def test_config4_ensemble_member_chunk(vfo):
    """BASELINE configs[3] (1,024 members x 8,192 profiles, X-mode, n_points = 20000) is streamed member by
    member; one member chunk of 1,024 perturbed profiles is run here and spot-checked against the oracle."""
    import torch
    lat, lon = synth.grid_subset(1024)
    alt = synth.default_alt()
    freq = synth.default_freq()
    den, bmag, bpsi = synth.ensemble_member(lat, lon, member=3, alt=alt)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    vh, st = vfo.vertical_forward_operator_batched(*t, 'X', 20000, return_status=True)
    assert int(st.abs().sum().item()) == 0
    a = vh.cpu().numpy()
    for p in (0, 511, 1023):
        ref = vfo_oracle.vertical_forward_operator(freq, den[p], bmag[p], bpsi[p], alt, 'X', 20000)
        assert np.array_equal(np.isnan(a[p]), np.isnan(ref)) and rel_err(a[p], ref) < 1e-9


@pytest.mark.parametrize("mode", ["O", "X"])
def test_solo_kernel_matches_two_kernel_modes(vfo, golden, mode, monkeypatch):
    """Single-profile calls run as ONE kernel (row setup + tile in the same CTA).  PRHF_NO_SOLO=1 forces the
    planned two-kernel mode; both must agree with each other and with the goldens."""
    from pyrayhf_b200 import _cabi
    fx = golden.fixtures
    for which in ("Day", "Night"):
        args = (fx["freq_a"], fx[which + "_den"], fx[which + "_bmag"], fx[which + "_bpsi"], fx[which + "_alt"])
        monkeypatch.delenv("PRHF_NO_SOLO", raising=False)
        monkeypatch.setattr(_cabi, "_contexts", {})
        a = vfo.vertical_forward_operator(*args, mode, 20000)
        a2 = vfo.vertical_forward_operator(*args, mode, 20000)
        assert np.array_equal(a, a2, equal_nan=True)
        monkeypatch.setenv("PRHF_NO_SOLO", "1")
        monkeypatch.setattr(_cabi, "_contexts", {})
        b = vfo.vertical_forward_operator(*args, mode, 20000)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and rel_err(a, b) < 1e-12
        tag = "%s_%s_20000_a" % (which, mode)
        assert_parity(a, fx["ref_" + tag], fx["truth_" + tag], mode, "solo " + tag)
    # failing profiles through the solo kernel
    den, bmag, bpsi, alt = synth.single_day_profile()
    monkeypatch.delenv("PRHF_NO_SOLO", raising=False)
    monkeypatch.setattr(_cabi, "_contexts", {})
    neg = den.copy()
    neg[3] = -1.0
    with pytest.raises(ValueError):
        vfo.vertical_forward_operator(fx["freq_a"], neg, bmag, bpsi, alt, mode, 20000)
    k = int(np.argmax(den))
    with pytest.raises(IndexError):
        vfo.vertical_forward_operator(fx["freq_a"], den[k:], bmag[k:], bpsi[k:], alt[k:], mode, 20000)

"""GPU: the standalone stage operators (pyrayhf_b200/stages.py -> vfo_stages.cu through the C ABI) against
the golden vectors captured from the live reference (tests/golden/stages.npz) and against the oracle.

Tolerances: the elementwise stages, the stretched altitudes, their spacings and the interpolated profile
restate the reference operation by operation -> 1e-13 relative (the only non-IEEE operation is exp in the
grid, <= 2 ulp); virtual heights 1e-9 (X-mode, isotropic) / the O-mode rounding-ball rule of conftest.
"""
import os
import warnings

import numpy as np
import pytest

from oracle import vfo_oracle
from pyrayhf_b200 import synth

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stages.npz")
REGRID_KEYS = ('freq', 'den', 'bmag', 'bpsi', 'dist', 'alt', 'crit_height', 'ind')


@pytest.fixture(scope="module")
def st():
    import torch
    assert torch.cuda.is_available()
    from pyrayhf_b200 import stages
    return stages


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def close(a, b, rtol=1e-13, atol=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.allclose(a[m], b[m], rtol=rtol, atol=atol), float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), 1e-300)))


def test_constants(st):
    assert st.constants() == (8.97866275, 2.799249247e10, 6371., 299_792.458)


def test_elementwise_bit_exact(st, g):
    # IEEE sqrt / multiply / divide in the reference's order: bit-for-bit
    assert np.array_equal(st.den2freq(g['ew_den']), g['ew_den2freq'])
    assert np.array_equal(st.find_X(g['ew_den'], g['ew_f']), g['ew_X'])
    assert np.array_equal(st.find_X(g['ew_den'][1], 3.3e6), g['ew_X_scalar_f'])
    assert np.array_equal(st.find_Y(g['ew_f'], g['ew_b']), g['ew_Y'])
    assert np.array_equal(st.find_Y(g['ew_f'][:, 0], 4.1e-5), g['ew_Y_scalar_b'])
    s = st.den2freq(4.0e10)
    assert np.ndim(s) == 0 and s == np.sqrt(4.0e10) * 8.97866275


def test_elementwise_errors_and_shapes(st):
    with pytest.raises(ValueError, match="Density must be non-negative"):
        st.den2freq(np.array([1.0, -1.0, 2.0]))
    with pytest.raises(ValueError, match="Density must be non-negative"):
        st.find_X(np.array([[1.0, 2.0], [3.0, -4.0]]), 1e6)
    assert np.isnan(st.den2freq(np.array([np.nan]))[0])                     # NaN passes the check (lib:93)
    assert st.find_X(np.zeros((0, 3)), 1e6).shape == (0, 3)
    assert st.find_Y(np.ones((2, 1)), np.ones(5)).shape == (2, 5)
    big = np.linspace(1e9, 1e12, 300001)
    assert np.array_equal(st.find_X(big, 5e6), vfo_oracle.plasma_ratio_x(big, 5e6))


def test_smooth_grids(st, g):
    for k, (s, e, n, sh) in enumerate(g['grid_args']):
        got = st.smooth_nonuniform_grid(s, e, int(n), sh)
        close(got, g['grid_%d' % k], rtol=1e-14, atol=1e-15)
    m = st.smooth_nonuniform_grid(0, 1, 20000, 10.)
    assert m[0] == 0.0 and m[-1] == 1.0 and np.all(np.diff(m) > 0)
    assert st.smooth_nonuniform_grid(0, 1, 0, 10.).shape == (0,)


def test_regrid_against_golden(st, g):
    for name in g['regrid_cases']:
        name = str(name)
        mode, n = str(g[name + '_mode']), int(g[name + '_n'])
        args = [g['%s_in_%s' % (name, k)] for k in ('f', 'den', 'bmag', 'bpsi', 'alt')]
        out = st.regrid_to_nonuniform_grid(*args, mode=mode, n_points=n)
        assert tuple(out.keys()) == REGRID_KEYS
        for k in REGRID_KEYS:
            ref = g['%s_out_%s' % (name, k)]
            assert out[k].shape == ref.shape, (name, k)
            if k == 'ind':
                assert np.array_equal(out[k], ref) and out[k].dtype.kind == 'i'
            elif k == 'dist':
                # differences of neighbouring altitudes: absolute error of 2 ulp of the altitude
                close(out[k], ref, rtol=1e-9, atol=1e-12)
            else:
                close(out[k], ref, rtol=1e-13)


def test_regrid_then_find_vh_equals_fused_operator(st, g):
    import pyrayhf_b200
    for name in g['regrid_cases']:
        name = str(name)
        mode, n = str(g[name + '_mode']), int(g[name + '_n'])
        f, den, bmag, bpsi, alt = (g['%s_in_%s' % (name, k)] for k in ('f', 'den', 'bmag', 'bpsi', 'alt'))
        out = st.regrid_to_nonuniform_grid(f, den, bmag, bpsi, alt, mode=mode, n_points=n)
        X = st.find_X(out['den'], out['freq'])
        Y = st.find_Y(out['freq'], out['bmag'])
        vh = st.find_vh(X, Y, out['bpsi'], out['dist'], np.min(alt), mode)
        ref = g[name + '_vh']
        assert np.array_equal(np.isnan(vh), np.isnan(ref)), name
        fused = pyrayhf_b200.vertical_forward_operator(f / 1e6, den, bmag, bpsi, alt, mode, n)
        assert np.array_equal(np.isnan(vh), np.isnan(fused)), name
        m = np.isfinite(ref)
        if mode == 'X' or name.startswith('day_B0'):
            assert np.allclose(vh[m], ref[m], rtol=1e-9, atol=0), name
        else:
            # O-mode: the numpy reference carries cancellation noise; the staged and the
            # fused GPU paths evaluate the same cancellation-free form and agree with each other
            assert np.allclose(vh[m], ref[m], rtol=5e-4, atol=0), name
        assert np.allclose(vh[m], fused[m], rtol=1e-9, atol=0), name
        # the reference's operation order (PRHF_FLAG_LITERAL): O-mode stays inside the cancellation noise, which
        # numpy's SIMD pow / sin / cos (<= 2 ulp from libdevice's) move around
        vh_lit = st.find_vh(g[name + '_X'], g[name + '_Y'], g['%s_out_bpsi' % name], g['%s_out_dist' % name],
                            np.min(alt), mode, literal=True)
        assert np.array_equal(np.isnan(vh_lit), np.isnan(ref)), name
        lit_tol = 1e-9 if (mode == 'X' or name.startswith('day_B0')) else 5e-4
        assert np.allclose(vh_lit[m], ref[m], rtol=lit_tol, atol=0), name


def test_regrid_full_size_against_oracle(st):
    den, bmag, bpsi, alt = synth.bench_day_profile()
    f = synth.default_freq() * 1e6
    out = st.regrid_to_nonuniform_grid(f, den, bmag, bpsi, alt, mode='X', n_points=20000,
                                       keys=('alt', 'den', 'bpsi', 'crit_height'))
    assert set(out) == {'alt', 'den', 'bpsi', 'crit_height'}
    ref = vfo_oracle.regrid(f, den, bmag, bpsi, alt, 'X', 20000)
    close(out['alt'], ref['h'], rtol=1e-13)
    close(out['den'], ref['den'], rtol=1e-12)
    close(out['bpsi'], ref['bpsi'], rtol=1e-12)
    close(out['crit_height'][:, 0], ref['h_c'], rtol=1e-15)


def test_regrid_errors(st):
    den, bmag, bpsi, alt = synth.single_day_profile()
    f = np.array([3e6, 5e6])
    with pytest.raises(ValueError, match="mode must be 'O' or 'X'"):
        st.regrid_to_nonuniform_grid(f, den, bmag, bpsi, alt, mode='o')
    bad = den.copy()
    bad[3] = -1.0
    with pytest.raises(ValueError, match="Density must be non-negative"):
        st.regrid_to_nonuniform_grid(f, bad, bmag, bpsi, alt)
    k = int(np.argmax(den))
    with pytest.raises(IndexError):                                          # density peak is the first sample
        st.regrid_to_nonuniform_grid(f, den[k:], bmag[k:], bpsi[k:], alt[k:])
    with pytest.raises(KeyError):
        st.regrid_to_nonuniform_grid(f, den, bmag, bpsi, alt, keys=('nope',))
    # the dh argument is ignored, as in the reference (lib:378)
    a = st.regrid_to_nonuniform_grid(f, den, bmag, bpsi, alt, n_points=20, dh=5.0)
    assert np.all(a['dist'][:, -1] == 1e-6)


def test_find_vh_mode_and_iso_switch(st):
    X = np.array([[0.1, 0.5, 0.9], [0.2, 0.4, 1.2]])
    dh = np.full_like(X, 2.0)
    psi = np.full_like(X, 40.0)
    Y0 = np.zeros_like(X)
    got = st.find_vh(X, Y0, psi, dh, 80.0, 'Q')                              # unmagnetised: mode never inspected
    ref = vfo_oracle.find_vh_rows(X, Y0, psi, dh, 80.0, 'Q')
    assert np.allclose(got, ref, rtol=1e-14)
    with pytest.raises(ValueError, match="Mode must be O or X"):
        st.find_vh(X, Y0 + 0.3, psi, dh, 80.0, 'Q')
    # one large |Y| anywhere switches the WHOLE array to the magnetised branch (lib:201)
    Y = Y0.copy()
    Y[1, 2] = 0.4
    got = st.find_vh(X, Y, psi, dh, 80.0, 'X')
    ref = vfo_oracle.find_vh_rows(X, Y, psi, dh, 80.0, 'X')
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.allclose(got, ref, rtol=1e-9, equal_nan=True)
    # all-NaN Y: nanmax is NaN -> magnetised branch -> every term NaN -> row sums 0 -> NaN
    got = st.find_vh(X, np.full_like(X, np.nan), psi, dh, 80.0, 'O')
    assert np.all(np.isnan(got))
    # long rows take the CTA-per-row path
    rng = np.random.default_rng(5)
    Xl = rng.uniform(0.0, 0.95, size=(5, 3000))
    Yl = rng.uniform(0.05, 0.4, size=(5, 3000))
    pl = rng.uniform(0.0, 90.0, size=(5, 3000))
    dl = rng.uniform(0.01, 1.0, size=(5, 3000))
    for mode in 'OX':
        got = st.find_vh(Xl, Yl, pl, dl, 90.0, mode)
        ref = vfo_oracle.find_vh_rows(Xl, Yl, pl, dl, 90.0, mode)
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.allclose(got, ref, rtol=1e-9, equal_nan=True)


def test_device_profile_generator_matches_host_generator():
    import torch
    import pyrayhf_b200
    lat, lon = synth.grid_subset(300)
    par = synth.ensemble_member_parameters(lat, lon, member=7)
    den_h, bmag_h, bpsi_h = synth.profiles_from_parameters(*par)
    den_d, bmag_d, bpsi_d = synth.profiles_from_parameters_device(*par)
    assert den_d.is_cuda and den_d.shape == den_h.shape
    close(den_d.cpu().numpy(), den_h, rtol=1e-11)
    close(bmag_d.cpu().numpy(), bmag_h, rtol=1e-13)
    close(bpsi_d.cpu().numpy(), bpsi_h, rtol=1e-12, atol=1e-12)
    # the forward operator on the device-built batch == on the host-built batch (no host copy of the profiles)
    dev = den_d.device
    freq = torch.from_numpy(synth.default_freq()).to(dev)
    alt = torch.from_numpy(synth.default_alt()).to(dev)
    vh_d = pyrayhf_b200.vertical_forward_operator_batched(freq, den_d, bmag_d, bpsi_d, alt, 'X', 2000).cpu().numpy()
    vh_h = pyrayhf_b200.vertical_forward_operator_batched(synth.default_freq(), den_h, bmag_h, bpsi_h,
                                                          synth.default_alt(), 'X', 2000)
    assert np.array_equal(np.isnan(vh_d), np.isnan(vh_h))
    m = np.isfinite(vh_h)
    assert np.allclose(vh_d[m], vh_h[m], rtol=1e-8, atol=0)     # inputs differ by ~1e-13; rows near cutoff amplify that

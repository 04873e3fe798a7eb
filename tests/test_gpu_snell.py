"""GPU: the batched Snell's-law tracers (pyrayhf_b200/snell.py -> vfo_snell.cu through the C ABI) against the
goldens captured from the live reference (tests/golden/snell.npz) and against the oracle.

Tolerance: 1e-9 relative on group path, group delay, ground range and the path coordinates for both tracers
(measured worst case on the goldens: 4.5e-14 flat Earth, 2.1e-12 spherical -- the spherical tracer's adaptive
midpoint rule evaluates p / (r sqrt((mu r)^2 - p^2)) up to 400 times per level next to the apex, where the
difference cancels, so it is the more sensitive one to the last bit of mu).  NaN masks (no ray) and the
number of path points must match exactly.  The midpoint / apex outputs must equal one of the reference's two
admissible nodes (oracle/snell_oracle.py explains why the reference itself is rounding-dependent there).
"""
import os
import warnings

import numpy as np
import pytest

from oracle import snell_oracle

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "snell.npz")
TOL = {'cartesian': 1e-9, 'spherical': 1e-9}


@pytest.fixture(scope="module")
def sn():
    import torch
    assert torch.cuda.is_available()
    from pyrayhf_b200 import snell
    return snell


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def profile(g, name):
    return tuple(g['%s_%s' % (name, k)] for k in ('alt', 'ne', 'babs', 'bpsi'))


def check_batch(got, ref, tol, tag):
    """ref columns: path, delay, x_mid, z_mid, ground, x_lo, z_lo, x_hi, z_hi, n_path."""
    assert np.array_equal(got['n_path'], ref[:, 9].astype(np.int32)), tag
    worst = 0.0
    for col, key in ((0, 'group_path_km'), (1, 'group_delay_sec'), (4, 'ground_range_km')):
        a, b = got[key], ref[:, col]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (tag, key)
        m = ~np.isnan(b)
        if m.any():
            err = np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))
            worst = max(worst, err)
            assert err <= tol, (tag, key, err)
    xm, zm = got['x_midpoint'], got['z_midpoint']
    assert np.array_equal(np.isnan(xm), np.isnan(ref[:, 2])), tag
    m = ~np.isnan(ref[:, 2])
    scale = np.maximum(np.abs(ref[m, 7]), 1.0)
    lo = (np.abs(xm[m] - ref[m, 5]) <= 10 * tol * scale) & (np.abs(zm[m] - ref[m, 6]) <= 10 * tol * scale)
    hi = (np.abs(xm[m] - ref[m, 7]) <= 10 * tol * scale) & (np.abs(zm[m] - ref[m, 8]) <= 10 * tol * scale)
    assert np.all(lo | hi), (tag, 'midpoint')
    assert np.array_equal(got['x_apex_km'], got['x_midpoint'], equal_nan=True)      # lib:1267-1268
    assert np.array_equal(got['z_apex_km'], got['z_midpoint'], equal_nan=True)
    return worst


@pytest.mark.parametrize("geo", ['cartesian', 'spherical'])
def test_batched_rays_against_golden(sn, g, geo):
    worst = 0.0
    for pn in g['profiles']:
        pn = str(pn)
        for mode in 'OX':
            tag = '%s_%s_%s' % (pn, geo, mode)
            got = sn.trace_rays_snells_batched(g[tag + '_f'], g[tag + '_el'], *profile(g, pn), mode, geometry=geo)
            worst = max(worst, check_batch(got, g[tag + '_out'], TOL[geo], tag))
    print("worst relative error", geo, worst)


def test_single_ray_dicts_and_paths(sn, g):
    for key in [k for k in g if k.startswith('path_') and k.endswith('_x')]:
        _, pn, geo, mode, el = key[:-2].split('_')
        fn = sn.trace_ray_cartesian_snells if geo == 'cartesian' else sn.trace_ray_spherical_snells
        r = fn(8e6, float(el), *profile(g, pn), mode)
        assert set(r) == set(snell_oracle.KEYS)
        assert r['x'].shape == g[key].shape
        assert np.allclose(r['z'], g[key[:-2] + '_z'], rtol=1e-12, atol=0)
        assert np.allclose(r['x'], g[key], rtol=TOL[geo], atol=1e-9)
        assert isinstance(r['group_path_km'], float)
    al, ne, bb, ps = profile(g, 'gauss')
    # no reflection: every key NaN; the spherical early exit has no apex keys (lib:1578-1585)
    r = sn.trace_ray_cartesian_snells(30e6, 60.0, al, ne, bb, ps, 'O')
    assert set(r) == set(snell_oracle.KEYS) and all(np.isnan(v) for v in r.values())
    r = sn.trace_ray_spherical_snells(30e6, 60.0, al, ne, bb, ps, 'O')
    assert set(r) == set(snell_oracle.KEYS[:7]) and all(np.isnan(v) for v in r.values())
    r = sn.trace_ray_spherical_snells(9e6, 35.0, al, ne, bb, ps, 'O', dz_target_km=0.25, apex_boost=50.0,
                                      max_substeps=1000, R_E=6371e3)
    got = np.array([r[k] for k in ("group_path_km", "group_delay_sec", "ground_range_km")])
    assert np.allclose(got, g['kw_out'][[0, 1, 4]], rtol=TOL['spherical'])


def test_fan_against_oracle_and_errors(sn, g):
    al, ne, bb, ps = profile(g, 'day')
    f = np.repeat(np.arange(3e6, 13e6, 1e6), 8)
    el = np.tile(np.linspace(10.0, 85.0, 8), 10)
    for geo in ('cartesian', 'spherical'):
        got = sn.trace_rays_snells_batched(f, el, al, ne, bb, ps, 'X', geometry=geo, return_paths=True)
        assert got['x'].shape == (80, 2 * (al.size + 1) + 1)
        for i in range(0, 80, 7):
            o = snell_oracle.trace(f[i], el[i], al, ne, bb, ps, 'X', geo)
            if np.ndim(o['x']) == 0:
                assert got['n_path'][i] == 0 and np.isnan(got['group_path_km'][i])
                continue
            n = got['n_path'][i]
            assert n == o['x'].size
            assert np.allclose(got['z'][i, :n], o['z'], rtol=1e-12)
            assert np.allclose(got['x'][i, :n], o['x'], rtol=TOL[geo], atol=1e-9)
            assert np.all(np.isnan(got['x'][i, n:]))
            assert abs(got['group_delay_sec'][i] - o['group_delay_sec']) <= TOL[geo] * o['group_delay_sec']
    with pytest.raises(ValueError, match="Mode must be O or X"):
        sn.trace_rays_snells_batched(5e6, 45.0, al, ne, bb, ps, 'Z')
    with pytest.raises(ValueError, match="Density must be non-negative"):
        sn.trace_rays_snells_batched(5e6, 45.0, al, -ne, bb, ps, 'O')
    with pytest.raises(ValueError):
        sn.trace_rays_snells_batched(5e6, 45.0, al, ne, bb, ps, 'O', geometry='toroidal')
    # literal evaluation order of the refractive index gives the same rays
    a = sn.trace_rays_snells_batched(f, el, al, ne, bb, ps, 'O')
    b = sn.trace_rays_snells_batched(f, el, al, ne, bb, ps, 'O', literal=True)
    assert np.array_equal(a['n_path'], b['n_path'])
    assert np.allclose(a['group_path_km'], b['group_path_km'], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("geometry", ["cartesian", "spherical"])
@pytest.mark.parametrize("mode", ["O", "X"])
def test_fan_entry_equals_the_per_ray_entry(geometry, mode):
    """prhf_snell_fan_f64 (field once per frequency) against prhf_snell_f64 on the written-out (frequency, elevation)
    pairs: every output bit for bit, paths included; the Python wrapper takes the fan entry by itself for inputs of
    shape [F, 1] x [E]."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.single_day_profile()
    f = np.linspace(1.0e6, 12.0e6, 23)
    e = np.linspace(3.0, 89.0, 17)
    fan = prhf.trace_rays_snells_batched(f[:, None], e[None, :], alt, den, bmag, bpsi, mode, geometry=geometry,
                                         return_paths=True)
    pairs = prhf.trace_rays_snells_batched(np.repeat(f, e.size), np.tile(e, f.size), alt, den, bmag, bpsi, mode,
                                           geometry=geometry, return_paths=True)
    assert fan["group_path_km"].shape == (f.size * e.size,)
    assert (fan["n_path"] > 0).sum() > 50
    for key in pairs:
        assert np.array_equal(fan[key], pairs[key], equal_nan=True), key

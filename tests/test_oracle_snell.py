"""CPU: the Snell-tracer oracle (oracle/snell_oracle.py) against the goldens captured from the live reference
(tests/golden/snell.npz) and, where /root/reference is mounted, against the reference itself."""
import os
import warnings

import numpy as np
import pytest

from oracle import snell_oracle
from oracle.ref_import import reference_available

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "snell.npz")
warnings.simplefilter("ignore")


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def profile(g, name):
    return tuple(g['%s_%s' % (name, k)] for k in ('alt', 'ne', 'babs', 'bpsi'))


def test_oracle_matches_golden_rays(g):
    n_checked = 0
    for pn in g['profiles']:
        pn = str(pn)
        al, ne, bb, ps = profile(g, pn)
        for geo in ('cartesian', 'spherical'):
            for mode in 'OX':
                tag = '%s_%s_%s' % (pn, geo, mode)
                ref = g[tag + '_out']
                for i in range(0, ref.shape[0], 3):                         # every third ray keeps the CPU suite short
                    o = snell_oracle.trace(g[tag + '_f'][i], g[tag + '_el'][i], al, ne, bb, ps, mode, geo)
                    got = np.array([o[k] for k in ("group_path_km", "group_delay_sec", "x_midpoint", "z_midpoint",
                                                   "ground_range_km")], dtype=float)
                    assert np.array_equal(np.isnan(got), np.isnan(ref[i, :5])), (tag, i)
                    m = ~np.isnan(got)
                    assert np.allclose(got[m], ref[i, :5][m], rtol=1e-12, atol=0), (tag, i)
                    assert np.atleast_1d(o['x']).size == (int(ref[i, 9]) or 1)
                    n_checked += 1
    assert n_checked > 200


def test_oracle_paths_and_keywords(g):
    al, ne, bb, ps = profile(g, 'gauss')
    for key in [k for k in g if k.startswith('path_') and k.endswith('_x')]:
        _, pn, geo, mode, el = key[:-2].split('_')
        o = snell_oracle.trace(8e6, float(el), *profile(g, pn), mode, geo)
        assert np.allclose(o['x'], g[key], rtol=1e-12, atol=1e-12)
        assert np.allclose(o['z'], g[key[:-2] + '_z'], rtol=1e-12, atol=0)
    o = snell_oracle.trace(9e6, 35.0, al, ne, bb, ps, 'O', 'spherical', dz_target_km=0.25, apex_boost=50.0,
                           max_substeps=1000, r_e=6371e3)
    got = np.array([o[k] for k in ("group_path_km", "group_delay_sec", "x_midpoint", "z_midpoint",
                                   "ground_range_km")])
    assert np.allclose(got, g['kw_out'], rtol=1e-12)


def test_midpoint_is_the_apex_or_the_node_below_it(g):
    """The rounding-dependent output of the reference: document it with the goldens themselves."""
    seen = set()
    for pn in g['profiles']:
        for geo in ('cartesian', 'spherical'):
            for mode in 'OX':
                ref = g['%s_%s_%s_out' % (str(pn), geo, mode)]
                ok = np.isfinite(ref[:, 2])
                lo = np.isclose(ref[ok, 2], ref[ok, 5], rtol=1e-12) & np.isclose(ref[ok, 3], ref[ok, 6], rtol=1e-12)
                hi = np.isclose(ref[ok, 2], ref[ok, 7], rtol=1e-12) & np.isclose(ref[ok, 3], ref[ok, 8], rtol=1e-12)
                assert np.all(lo | hi)
                seen.update(np.where(lo, 'below', 'apex').tolist())
    assert seen == {'below', 'apex'}


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference():
    from oracle.ref_import import load_reference_library
    lib = load_reference_library()
    alt = np.linspace(0, 600, 200)
    ne = 1e12 * np.exp(-(alt - 250) ** 2 / (2 * 60 ** 2))
    bb, ps = np.full_like(alt, 4e-5), np.full_like(alt, 45.0)
    for f, el, mode in ((10e6, 45.0, 'O'), (7e6, 30.0, 'X'), (12e6, 75.0, 'O'), (30e6, 60.0, 'O')):
        for geo, fn in (('cartesian', lib.trace_ray_cartesian_snells), ('spherical', lib.trace_ray_spherical_snells)):
            r = fn(f, el, alt, ne, bb, ps, mode)
            o = snell_oracle.trace(f, el, alt, ne, bb, ps, mode, geo)
            assert set(r) == {k for k in o if not k.startswith('mid_')}
            for k in r:
                assert np.allclose(np.asarray(r[k], float), np.asarray(o[k], float), rtol=1e-12, atol=1e-12,
                                   equal_nan=True), (geo, k)

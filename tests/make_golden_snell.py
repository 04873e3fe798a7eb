"""Generate tests/golden/snell.npz from the LIVE reference (dev container only).

    python tests/make_golden_snell.py

Rays traced by /root/reference/PyRayHF/library.py trace_ray_cartesian_snells / trace_ray_spherical_snells
(imported through oracle/ref_import.py) over three profiles: the Gaussian layer of the reference's own tests
(tests/test_core.py:724-733), the tutorial Day profile and the synthetic Chapman day profile.  Stored per ray:
the five scalars, the number of path points, and the two admissible midpoints (apex / node below the apex: which
one the reference's searchsorted returns depends on the last bit of a cumulative sum, see oracle/snell_oracle.py).
Full paths are stored for a few rays.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference_library, load_tutorial_fixture  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.simplefilter("ignore")
SCALARS = ("group_path_km", "group_delay_sec", "x_midpoint", "z_midpoint", "ground_range_km")


def main():
    lib = load_reference_library()
    g = {}
    alt = np.linspace(0, 600, 200)
    profs = {'gauss': (alt, 1e12 * np.exp(-(alt - 250) ** 2 / (2 * 60 ** 2)), np.full_like(alt, 4e-5),
                       np.full_like(alt, 45.0))}
    day = load_tutorial_fixture('Day')
    profs['day'] = tuple(np.asarray(day[k], dtype=float) for k in ('alt', 'den', 'bmag', 'bpsi'))
    den, bmag, bpsi, salt = synth.single_day_profile()
    profs['synth'] = (salt, den, bmag, bpsi)
    profs['gauss_b0'] = (alt, profs['gauss'][1], np.zeros_like(alt), profs['gauss'][3])
    g['profiles'] = np.array(sorted(profs))
    freqs = (2e6, 5e6, 8e6, 10e6, 14e6, 25e6)
    elevs = (5., 20., 45., 60., 80., 89.9, 90.)
    for pn, (al, ne, bb, ps) in profs.items():
        for k, v in (('alt', al), ('ne', ne), ('babs', bb), ('bpsi', ps)):
            g['%s_%s' % (pn, k)] = v
        for geo in ('cartesian', 'spherical'):
            fn = lib.trace_ray_cartesian_snells if geo == 'cartesian' else lib.trace_ray_spherical_snells
            for mode in 'OX':
                rows, f_list, e_list = [], [], []
                for f in freqs:
                    for el in elevs:
                        r = fn(f, el, al, ne, bb, ps, mode)
                        x, z = np.atleast_1d(r['x']), np.atleast_1d(r['z'])
                        if x.size > 1 or np.isfinite(r['group_path_km']):
                            n_up = (x.size + 1) // 2
                            lo = max(n_up - 2, 0)
                            cand = [x[lo], z[lo], x[n_up - 1], z[n_up - 1]]
                            n_path = x.size
                        else:
                            cand, n_path = [np.nan] * 4, 0
                        rows.append([r[k] for k in SCALARS] + cand + [n_path])
                        f_list.append(f)
                        e_list.append(el)
                        if pn in ('gauss', 'day') and f == 8e6 and el in (45., 80.):
                            tag = 'path_%s_%s_%s_%d' % (pn, geo, mode, int(el))
                            g[tag + '_x'], g[tag + '_z'] = x, z
                tag = '%s_%s_%s' % (pn, geo, mode)
                g[tag + '_f'] = np.array(f_list)
                g[tag + '_el'] = np.array(e_list)
                g[tag + '_out'] = np.array(rows, dtype=float)
                fin = np.isfinite(np.array(rows, dtype=float)[:, 0]).sum()
                print(tag, 'rays', len(rows), 'with a path', int(fin))
    # keyword controls of the spherical tracer
    al, ne, bb, ps = profs['gauss']
    r = lib.trace_ray_spherical_snells(9e6, 35.0, al, ne, bb, ps, 'O', dz_target_km=0.25, apex_boost=50.0,
                                       max_substeps=1000, R_E=6371e3)
    g['kw_out'] = np.array([r[k] for k in SCALARS], dtype=float)
    np.savez_compressed(os.path.join(GOLDEN, "snell.npz"), **g)
    print("written", os.path.getsize(os.path.join(GOLDEN, "snell.npz")), "bytes")


if __name__ == "__main__":
    main()

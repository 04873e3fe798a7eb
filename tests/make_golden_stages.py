"""Generate tests/golden/stages.npz from the LIVE reference (dev container only).

    python tests/make_golden_stages.py

Golden vectors for the standalone stages of the path (pyrayhf_b200/stages.py): den2freq, find_X, find_Y,
smooth_nonuniform_grid, regrid_to_nonuniform_grid and find_vh of /root/reference/PyRayHF/library.py, imported
through oracle/ref_import.py.  Inputs and the reference's outputs are stored side by side.
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

REGRID_KEYS = ('freq', 'den', 'bmag', 'bpsi', 'dist', 'alt', 'crit_height', 'ind')


def main():
    lib = load_reference_library()
    rng = np.random.default_rng(77)
    g = {}
    # ---- elementwise stages ----
    den = 10.0 ** rng.uniform(8, 12.5, size=(7, 33))
    den[0, 0] = 0.0
    f = rng.uniform(0.1e6, 20e6, size=(7, 1))
    b = rng.uniform(2e-5, 6e-5, size=33)
    g['ew_den'], g['ew_f'], g['ew_b'] = den, f, b
    g['ew_den2freq'] = lib.den2freq(den)
    g['ew_X'] = lib.find_X(den, f)
    g['ew_X_scalar_f'] = lib.find_X(den[1], 3.3e6)
    g['ew_Y'] = lib.find_Y(f, b)
    g['ew_Y_scalar_b'] = lib.find_Y(f[:, 0], 4.1e-5)
    # ---- smooth grids: (start, end, n, sharpness) ----
    grids = [(0, 1, 200, 10.), (0, 1, 2000, 10.), (0, 1, 1, 10.), (0, 1, 2, 10.), (2., 5., 37, 3.5), (0., 1., 50, 0.5)]
    g['grid_args'] = np.array(grids, dtype=np.float64)
    for k, (s, e, n, sh) in enumerate(grids):
        g['grid_%d' % k] = lib.smooth_nonuniform_grid(s, e, n, sh)
    # ---- regrid + find_vh ----
    cases = {}
    day = load_tutorial_fixture("Day")
    night = load_tutorial_fixture("Night")
    fsub = np.arange(0.5, 17.5, 0.5) * 1e6
    cases['day_O_64'] = (fsub, day, 'O', 64)
    cases['day_X_64'] = (fsub, day, 'X', 64)
    cases['night_X_50'] = (np.arange(1.0, 12.0, 0.7) * 1e6, night, 'X', 50)
    sden, sbmag, sbpsi, salt = synth.single_day_profile()
    cases['synth_O_33'] = (np.array([0.05, 1.1, 2.5, 4.0, 7.3, 9.9, 12.0, 30.0]) * 1e6,
                           dict(den=sden, bmag=sbmag, bpsi=sbpsi, alt=salt), 'O', 33)
    cases['synth_X_1'] = (np.array([2.5, 7.3, 30.0]) * 1e6, dict(den=sden, bmag=sbmag, bpsi=sbpsi, alt=salt), 'X', 1)
    # peak at index 1: one level survives the truncation (numpy's single-node interp quirk on dead rows)
    d1 = dict(den=np.array([1e11, 5e11, 2e11, 1e11]), bmag=np.full(4, 4e-5), bpsi=np.full(4, 30.0),
              alt=np.array([100., 110., 120., 130.]))
    cases['one_level_X_9'] = (np.array([1.0, 2.0, 3.5, 9.0]) * 1e6, d1, 'X', 9)
    # unmagnetised profile
    d0 = dict(den=np.asarray(day['den'], float), bmag=np.zeros_like(np.asarray(day['den'], float)),
              bpsi=np.asarray(day['bpsi'], float), alt=np.asarray(day['alt'], float))
    cases['day_B0_O_40'] = (fsub[::3], d0, 'O', 40)
    g['regrid_cases'] = np.array(sorted(cases))
    for name, (fq, d, mode, n) in cases.items():
        den_, bm_, ps_, al_ = (np.asarray(d[k], dtype=np.float64) for k in ('den', 'bmag', 'bpsi', 'alt'))
        out = lib.regrid_to_nonuniform_grid(fq, den_, bm_, ps_, al_, mode=mode, n_points=n)
        for k, v in (('f', fq), ('den', den_), ('bmag', bm_), ('bpsi', ps_), ('alt', al_)):
            g['%s_in_%s' % (name, k)] = v
        g[name + '_mode'] = np.array(mode)
        g[name + '_n'] = np.array(n)
        for k in REGRID_KEYS:
            g['%s_out_%s' % (name, k)] = np.ascontiguousarray(out[k])
        X = lib.find_X(out['den'], out['freq'])
        Y = lib.find_Y(out['freq'], out['bmag'])
        vh = lib.find_vh(X, Y, out['bpsi'], out['dist'], np.min(al_), mode)
        g[name + '_X'] = X
        g[name + '_Y'] = Y
        g[name + '_vh'] = vh
        full = lib.vertical_forward_operator(fq / 1e6, den_, bm_, ps_, al_, mode=mode, n_points=n)
        assert np.array_equal(np.isnan(full), np.isnan(vh)), name
        print(name, 'finite rows', int(np.isfinite(vh).sum()), 'of', vh.size)
    np.savez_compressed(os.path.join(GOLDEN, "stages.npz"), **g)
    print("written", os.path.join(GOLDEN, "stages.npz"), os.path.getsize(os.path.join(GOLDEN, "stages.npz")))


if __name__ == "__main__":
    main()

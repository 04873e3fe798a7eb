"""CPU: the oracle's stage restatements against the golden vectors captured from the live reference
(tests/make_golden_stages.py -> tests/golden/stages.npz)."""
import os
import warnings

import numpy as np
import pytest

from oracle import vfo_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stages.npz")
REGRID_KEYS = ('freq', 'den', 'bmag', 'bpsi', 'dist', 'alt', 'crit_height', 'ind')


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def close(a, b, rtol=1e-13):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    assert np.allclose(a[m], b[m], rtol=rtol, atol=0.0)


def test_elementwise(g):
    close(vfo_oracle.den2freq(g['ew_den']), g['ew_den2freq'])
    close(vfo_oracle.plasma_ratio_x(g['ew_den'], g['ew_f']), g['ew_X'])
    close(vfo_oracle.plasma_ratio_x(g['ew_den'][1], 3.3e6), g['ew_X_scalar_f'])
    close(vfo_oracle.gyro_ratio_y(g['ew_f'], g['ew_b']), g['ew_Y'])
    close(vfo_oracle.gyro_ratio_y(g['ew_f'][:, 0], 4.1e-5), g['ew_Y_scalar_b'])
    with pytest.raises(ValueError, match="Density must be non-negative"):
        vfo_oracle.den2freq(np.array([1.0, -1.0]))


def test_smooth_grids(g):
    for k, (s, e, n, sh) in enumerate(g['grid_args']):
        close(vfo_oracle.smooth_grid(s, e, int(n), sh), g['grid_%d' % k])
    close(vfo_oracle.stretch_multiplier(200), g['grid_0'])


def test_regrid_and_find_vh(g):
    for name in g['regrid_cases']:
        name = str(name)
        mode, n = str(g[name + '_mode']), int(g[name + '_n'])
        args = [g['%s_in_%s' % (name, k)] for k in ('f', 'den', 'bmag', 'bpsi', 'alt')]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = vfo_oracle.regrid_dict(*args, mode, n)
            for k in REGRID_KEYS:
                close(out[k], g['%s_out_%s' % (name, k)])
            vh = vfo_oracle.find_vh_rows(g[name + '_X'], g[name + '_Y'], out['bpsi'], out['dist'],
                                         np.min(args[4]), mode)
        close(vh, g[name + '_vh'], rtol=1e-12)

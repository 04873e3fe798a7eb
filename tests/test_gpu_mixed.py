"""Optional mixed single/double precision mode (PRHF_FLAG_MIXED_F32; BASELINE.json north_star "a documented bound for an
optional FP32 mode, including NaN / no-reflection masks matching exactly").  The bound asserted here is the one DESIGN.md
documents: 1e-5 relative on every finite virtual height (measured: ~1e-6 worst, ~1e-7 typical), masks identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BOUND = 1e-5


def _cmp(a, b, label):
    assert np.array_equal(np.isnan(a), np.isnan(b)), label
    m = np.isfinite(b)
    err = np.abs(a[m] - b[m]) / np.abs(b[m])
    assert err.max(initial=0.0) <= BOUND, (label, float(err.max()))
    return float(err.max(initial=0.0)), float(np.median(err)) if err.size else 0.0


@pytest.mark.parametrize("mode", ["X", "O"])
def test_mixed_mode_stays_within_its_documented_bound(golden, mode):
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    fx = golden.fixtures
    worst = 0.0
    for which in ("Day", "Night"):
        args = (fx["freq_a"], fx[which + "_den"], fx[which + "_bmag"], fx[which + "_bpsi"], fx[which + "_alt"])
        for n in (2048, 20000):
            f64 = prhf.vertical_forward_operator(*args, mode, n)
            mix = prhf.vertical_forward_operator(*args, mode, n, precision='mixed')
            worst = max(worst, _cmp(mix, f64, "%s %s n=%d" % (which, mode, n))[0])
            assert not np.array_equal(mix, f64, equal_nan=True)      # the flag does switch arithmetic
            # against the reference / truth as well: the mixed result is within the bound of the real thing
            ref = fx["truth_%s_%s_%d_a" % (which, mode, n)] if n == 20000 else None
            if ref is not None:
                _cmp(mix, ref, "%s %s vs truth" % (which, mode))
    lat, lon = synth.grid_subset(96, seed=21)
    alt, freq = synth.default_alt(), synth.default_freq()
    d, b, p = synth.profiles_at(lat, lon, alt)
    p = p + np.linspace(0.0, 2.0, alt.size)[None, :] * (np.arange(96)[:, None] % 2)   # every other profile: rotating angle
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, d, b, p, alt)]
    for n in (6000, 20000):
        f64 = prhf.vertical_forward_operator_batched(*t, mode, n, errors='nan').cpu().numpy()
        mix = prhf.vertical_forward_operator_batched(*t, mode, n, errors='nan', precision='mixed').cpu().numpy()
        worst = max(worst, _cmp(mix, f64, "batch %s n=%d" % (mode, n))[0])
    print("mixed mode %s: worst relative deviation from float64 %.2e" % (mode, worst))
    # small n_points (row-per-warp kernel) and literal: flag accepted, double precision used
    a = prhf.vertical_forward_operator_batched(*t, mode, 200, errors='nan', precision='mixed').cpu().numpy()
    b2 = prhf.vertical_forward_operator_batched(*t, mode, 200, errors='nan').cpu().numpy()
    assert np.array_equal(a, b2, equal_nan=True)
    with pytest.raises(ValueError):
        prhf.vertical_forward_operator_batched(*t, mode, 200, precision='float32')

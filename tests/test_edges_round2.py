"""Contract edges added in round 2 (VERDICT r1 "missing" 5-6): altitude grids that are not increasing,
profiles with more levels than the shared-memory staging holds, concurrent calls from two host threads.
Goldens: tests/golden/edge2.npz, generated from the live reference by tests/make_golden_edge2.py."""
import os
import threading

import numpy as np
import pytest

from conftest import assert_parity

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edge2.npz"))


def _day():
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.bench_day_profile()
    return synth.default_freq(), den, bmag, bpsi, alt


# ---------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize("mode", ["O", "X"])
def test_oracle_on_non_increasing_altitudes_is_the_reference(mode):
    from oracle import vfo_oracle
    freq, den, bmag, bpsi, alt = _day()
    for n in (200, 2000):
        got = vfo_oracle.vertical_forward_operator(freq, den[::-1].copy(), bmag[::-1].copy(), bpsi[::-1].copy(),
                                                   alt[::-1].copy(), mode, n)
        assert np.array_equal(got, GOLD["reversed_%s_%d" % (mode, n)], equal_nan=True)
        got = vfo_oracle.vertical_forward_operator(freq, den, bmag, bpsi, alt[::-1].copy(), mode, n)
        assert np.array_equal(got, GOLD["altdown_%s_%d" % (mode, n)], equal_nan=True)


def test_oracle_on_long_profiles_is_the_reference():
    from oracle import vfo_oracle
    for mode in ("O", "X"):
        got = vfo_oracle.vertical_forward_operator(GOLD["fsub"], GOLD["long5000_den"][1], GOLD["long5000_bmag"][1],
                                                   GOLD["long5000_bpsi"][1], GOLD["long5000_alt"], mode, 200)
        assert np.array_equal(got, GOLD["long5000_%s_200" % mode][1], equal_nan=True)


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["O", "X"])
def test_non_increasing_altitude_grid_matches_the_reference(mode):
    """np.interp on a decreasing axis: its range tests give every query the last truncated level (constant X, Y, psi),
    which leaves 0 (O) / 4-5 (X) finite rows of 174.  Single-profile entry, batched entry, both kernel families."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    freq, den, bmag, bpsi, alt = _day()
    rev = [np.ascontiguousarray(v[::-1]) for v in (den, bmag, bpsi, alt)]
    for n in (200, 2000):
        for key, args in (("reversed", (rev[0], rev[1], rev[2], rev[3])), ("altdown", (den, bmag, bpsi, rev[3]))):
            ref = GOLD["%s_%s_%d" % (key, mode, n)]
            got = prhf.vertical_forward_operator(freq, *args, mode, n)
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (key, mode, n)
            m = np.isfinite(ref)
            if m.any():
                np.testing.assert_allclose(got[m], ref[m], rtol=1e-9)
    # batched: a reversed profile between two ordinary ones, per-profile altitude grids
    dn, bn, pn = GOLD["night_den"], GOLD["night_bmag"], GOLD["night_bpsi"]
    d3 = np.stack([den, dn[::-1], den])
    b3 = np.stack([bmag, bn[::-1], bmag])
    p3 = np.stack([bpsi, pn[::-1], bpsi])
    a3 = np.stack([alt, alt[::-1], alt])
    got = prhf.vertical_forward_operator_batched(freq, d3, b3, p3, a3, mode, 200, errors='nan')
    ref1 = GOLD["night_reversed_%s_200" % mode]
    assert np.array_equal(np.isnan(got[1]), np.isnan(ref1))
    m = np.isfinite(ref1)
    if m.any():
        np.testing.assert_allclose(got[1][m], ref1[m], rtol=1e-9)
    single = prhf.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, 200)
    assert np.array_equal(got[0], got[2], equal_nan=True) and np.array_equal(np.isnan(got[0]), np.isnan(single))
    del synth


@pytest.mark.gpu
@pytest.mark.parametrize("n_alt", [3000, 5000])
@pytest.mark.parametrize("mode", ["O", "X"])
def test_profiles_longer_than_the_shared_memory_staging(n_alt, mode):
    """n_alt > prhf_max_n_alt(): the global-memory form of the operator (row setup reading the levels in place, node
    table in global memory).  Parity against the live-reference goldens and the long-double truth."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import _cabi
    from oracle import scalar, vfo_oracle
    assert n_alt > _cabi.context(-1).max_n_alt()
    alt, den = GOLD["long%d_alt" % n_alt], GOLD["long%d_den" % n_alt]
    bmag, bpsi, fsub = GOLD["long%d_bmag" % n_alt], GOLD["long%d_bpsi" % n_alt], GOLD["fsub"]
    for n in (200, 5000):
        ref = GOLD["long%d_%s_%d" % (n_alt, mode, n)]
        mult = vfo_oracle.stretch_multiplier(n)
        truth = scalar.vertical_forward_operator_batched(fsub, den, bmag, bpsi, alt, mode, n, variant=1, multiplier=mult)[0]
        got = prhf.vertical_forward_operator_batched(fsub, den, bmag, bpsi, alt, mode, n)
        assert_parity(got, ref, truth, mode, "long %d n=%d" % (n_alt, n))
        one = prhf.vertical_forward_operator(fsub, den[1], bmag[1], bpsi[1], alt, mode, n)
        assert np.array_equal(one, got[1], equal_nan=True)
    # literal flag and a failed profile go through the same form
    lit = prhf.vertical_forward_operator_batched(fsub, den, bmag, bpsi, alt, mode, 200, literal=True)
    assert np.array_equal(np.isnan(lit), np.isnan(GOLD["long%d_%s_200" % (n_alt, mode)]))
    bad = den.copy()
    bad[0, 10] = -1.0
    with pytest.raises(ValueError, match="Density must be non-negative"):
        prhf.vertical_forward_operator_batched(fsub, bad, bmag, bpsi, alt, mode, 200)


@pytest.mark.gpu
def test_two_host_threads_call_the_drop_in_concurrently():
    """Re-entrancy (the reference is a pure function): two threads hammer the numpy drop-in and the batched
    entry at once; every result must equal the serial one bit for bit."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    freq, den, bmag, bpsi, alt = _day()
    lat, lon = synth.grid_subset(6, seed=3)
    d6, b6, p6 = synth.profiles_at(lat, lon, alt)
    want = {("X", 2000): prhf.vertical_forward_operator(freq, den, bmag, bpsi, alt, "X", 2000),
            ("O", 300): prhf.vertical_forward_operator(freq, den, bmag, bpsi, alt, "O", 300)}
    want_b = prhf.vertical_forward_operator_batched(freq, d6, b6, p6, alt, "X", 2500, errors='nan')
    errors = []

    def worker(mode, n):
        try:
            for k in range(60):
                got = prhf.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n)
                if not np.array_equal(got, want[(mode, n)], equal_nan=True):
                    errors.append((mode, n, k, "single"))
                if k % 6 == 0:
                    gb = prhf.vertical_forward_operator_batched(freq, d6, b6, p6, alt, "X", 2500, errors='nan')
                    if not np.array_equal(gb, want_b, equal_nan=True):
                        errors.append((mode, n, k, "batched"))
        except Exception as exc:          # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=a) for a in (("X", 2000), ("O", 300))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


@pytest.mark.gpu
def test_batched_calls_on_two_torch_streams_are_ordered_on_the_device():
    """One ctx, two streams: the second call's stream waits for the first; both results are right."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    freq, den, bmag, bpsi, alt = _day()
    lat, lon = synth.grid_subset(40, seed=9)
    d, b, p = synth.profiles_at(lat, lon, alt)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, d, b, p, alt)]
    want1 = prhf.vertical_forward_operator_batched(*t, "X", 6000, errors='nan').cpu().numpy()
    want2 = prhf.vertical_forward_operator_batched(t[0], t[1][:5], t[2][:5], t[3][:5], t[4], "X", 3000, errors='nan').cpu().numpy()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    for _ in range(5):
        with torch.cuda.stream(s1):
            g1 = prhf.vertical_forward_operator_batched(*t, "X", 6000, errors='nan')
        with torch.cuda.stream(s2):
            g2 = prhf.vertical_forward_operator_batched(t[0], t[1][:5], t[2][:5], t[3][:5], t[4], "X", 3000, errors='nan')
        torch.cuda.synchronize()
        assert np.array_equal(g1.cpu().numpy(), want1, equal_nan=True)
        assert np.array_equal(g2.cpu().numpy(), want2, equal_nan=True)


@pytest.mark.gpu
def test_batched_shape_validation_and_broadcasts():
    """raw pointers only leave Python after every shape has been checked."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    freq, den, bmag, bpsi, alt = _day()
    lat, lon = synth.grid_subset(4, seed=2)
    d, b, p = synth.profiles_at(lat, lon, alt)
    full = prhf.vertical_forward_operator_batched(freq, d, np.tile(bmag, (4, 1)), np.tile(bpsi, (4, 1)), alt, "X", 300)
    shared = prhf.vertical_forward_operator_batched(freq, d, bmag, bpsi, alt, "X", 300)          # [A] broadcast
    assert np.array_equal(full, shared, equal_nan=True)
    dev = torch.device("cuda:0")
    tt = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, d, bmag, bpsi, alt)]
    assert np.array_equal(prhf.vertical_forward_operator_batched(*tt, "X", 300).cpu().numpy(), full, equal_nan=True)
    out = np.empty((4, freq.size))
    assert prhf.vertical_forward_operator_batched(freq, d, b, p, alt, "X", 300, out=out) is out
    for bad in (lambda: prhf.vertical_forward_operator_batched(freq, d, bmag[:-1], bpsi, alt, "X", 300),
                lambda: prhf.vertical_forward_operator_batched(freq, d, b[:3], p, alt, "X", 300),
                lambda: prhf.vertical_forward_operator_batched(freq, d, b, p, alt[:-1], "X", 300),
                lambda: prhf.vertical_forward_operator_batched(np.tile(freq, (3, 1)), d, b, p, alt, "X", 300),
                lambda: prhf.vertical_forward_operator_batched(freq, d[0], b[0], p[0], alt, "X", 300),
                lambda: prhf.vertical_forward_operator_batched(freq, d, b, p, alt, "X", 300, out=np.empty((4, 3))),
                lambda: prhf.vertical_forward_operator_batched(freq, d, b, p, alt, "X", 0),
                lambda: prhf.vertical_forward_operator_batched(*tt[:2], tt[2][:-1], *tt[3:], "X", 300),
                lambda: prhf.vertical_forward_operator_batched(*tt, "X", 300, out=torch.empty((4, 3), dtype=torch.float64, device=dev))):
        with pytest.raises(ValueError):
            bad()

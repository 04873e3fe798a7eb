"""Round-2 decompositions of the tile kernel: the E-space grid loop (the stretched grid as a geometric sequence, no table
reads), the live-row queue of large batches, and the rows the row setup finishes itself.

* odd / even n_points, tile boundaries and re-seed boundaries of the E-space loop against the scalar C restatement of
  library.py:459-509 (X-mode) and the long-double truth (O-mode) -- the last pair of points of a row is added outside
  the loop and differs between odd and even n_points;
* queued mode (default: narrow CTAs drawing whole rows by ticket) against one full-width tile-kernel CTA per row
  (PRHF_QUEUE=0) on a batch that has rows without reflection, rows clamped to the first level (finished by the row
  setup in queued mode) and ordinary rows; rows whose level window exceeds the narrow kernel's node buffer are deferred
  to the full-width kernel.
"""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu


def _truth_and_literal(freq, den, bmag, bpsi, alt, mode, n):
    from oracle import scalar, vfo_oracle
    mult = vfo_oracle.stretch_multiplier(n)
    if den.ndim == 1:
        den, bmag, bpsi = den[None], bmag[None], bpsi[None]
    lit = scalar.vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode, n, variant=0, multiplier=mult)[0]
    tru = scalar.vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode, n, variant=1, multiplier=mult)[0]
    return lit, tru


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n_points", [2048, 2049, 3001, 4097, 4098, 5001, 12345, 20001, 33281, 40000])
def test_espace_loop_odd_even_and_boundaries_single_profile(mode, n_points):
    """One profile (single-launch kernel, rows split into segments): the row's last pair of points, the tile boundaries
    (multiples of 512 points) and, from 32 769 points per tile on, the re-seed of the recurrence."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.bench_day_profile()
    freq = synth.default_freq()[::3]
    got = prhf.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points)
    lit, tru = _truth_and_literal(freq, den, bmag, bpsi, alt, mode, n_points)
    assert_parity(got, lit[0], tru[0], mode, "single profile n=%d %s" % (n_points, mode))


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("n_points,n_prof", [(4097, 5), (5000, 3), (9999, 7), (20000, 2), (33000, 24)])
def test_espace_loop_small_batches(mode, n_points, n_prof):
    """2 ... 24 profiles: planned mode (segments sized from the live-row count) and, from 24 profiles on, one CTA per
    row with the warp-per-frequency row setup."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    alt, freq = synth.default_alt(), synth.default_freq()[::5]
    lat, lon = synth.grid_subset(max(n_prof, 8))
    den, bmag, bpsi = synth.profiles_at(lat[:n_prof], lon[:n_prof], alt)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    got = prhf.vertical_forward_operator_batched(*t, mode, n_points).cpu().numpy()
    sample = sorted(set([0, n_prof // 2, n_prof - 1]))
    lit, tru = _truth_and_literal(freq, den[sample], bmag[sample], bpsi[sample], alt, mode, n_points)
    for k, q in enumerate(sample):
        assert_parity(got[q], lit[k], tru[k], mode, "batch %d x n=%d %s profile %d" % (n_prof, n_points, mode, q))


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("literal,n_points,f_step", [(False, 4500, 1), (True, 4500, 1), (False, 40000, 6)])
def test_queued_mode_equals_one_cta_per_row(mode, literal, n_points, f_step, monkeypatch):
    """>= 148 profiles and more than 4096 grid points per row: the row setup (one thread per frequency) queues the rows
    that reflect and finishes the rows clamped to the first level in closed form; PRHF_QUEUE=0 is the round-1 form
    with one full-width tile-kernel CTA per row.  Same masks and status, values to 1e-12 (128 against 256 threads per
    tile), clamped rows bit for bit; oracle on sampled profiles."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import _cabi, synth
    # (40 000 points per row: a whole-row tile is 79 iterations per thread, past the re-seed of the E recurrence)
    alt, freq = synth.default_alt(), synth.default_freq()[::f_step]
    lat, lon = synth.grid_subset(300)
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    den[7, :] = 0.0                                           # a profile without plasma: peak at index 0 -> status 2
    den[11, 5] = -1.0                                         # negative density: status 1
    dev = torch.device("cuda:0")

    def run():
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
        vh, st = prhf.vertical_forward_operator_batched(*t, mode, n_points, literal=literal, errors='nan',
                                                        return_status=True)
        return vh.cpu().numpy(), st.cpu().numpy()

    a, sa = run()
    monkeypatch.setenv("PRHF_QUEUE", "0")
    monkeypatch.setattr(_cabi, "_contexts", {})
    b, sb = run()
    assert np.array_equal(sa, sb) and sa[7] == 2 and sa[11] == 1 and np.count_nonzero(sa) == 2
    # same rows, other thread count per tile (128 against 256): the sums associate differently
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(a)
    assert np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) < 1e-12
    # rows clamped to the first level are finished by the row setup in queued mode and by the tile kernel otherwise:
    # the same closed form, the same bits
    clamped = m & (np.abs(a - alt.min()) < 1e-9)
    assert np.array_equal(a[clamped], b[clamped])
    assert np.isnan(a[7]).all() and np.isnan(a[11]).all()
    if not literal:
        sample = [0, 150, 299]
        lit, tru = _truth_and_literal(freq, den[sample], bmag[sample], bpsi[sample], alt, mode, n_points)
        for k, q in enumerate(sample):
            assert_parity(a[q], lit[k], tru[k], mode, "queued %s profile %d" % (mode, q))


@pytest.mark.parametrize("mode", ["O", "X"])
def test_queued_mode_defers_rows_whose_window_exceeds_the_node_buffer(mode, monkeypatch):
    """A 1 500-level altitude grid (0.41 km spacing): the narrow queue kernel holds 434 levels per CTA on B200, so every
    row that reflects above ~260 km is deferred to the full-width kernel that follows; rows below it stay.  Against
    the one-CTA-per-row form and the oracle."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import _cabi, synth
    alt = np.linspace(80.0, 699.0, 1500)
    freq = synth.default_freq()[::2]
    lat, lon = synth.grid_subset(160)
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    assert den.argmax(axis=1).max() > 470                    # windows beyond the narrow kernel's buffer exist
    n_points = 6000
    dev = torch.device("cuda:0")

    def run():
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
        return prhf.vertical_forward_operator_batched(*t, mode, n_points).cpu().numpy()

    a = run()
    monkeypatch.setenv("PRHF_QUEUE", "0")
    monkeypatch.setattr(_cabi, "_contexts", {})
    b = run()
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(a)
    assert m.sum() > 1000 and np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) < 1e-12
    sample = [0, 80, 159]
    lit, tru = _truth_and_literal(freq, den[sample], bmag[sample], bpsi[sample], alt, mode, n_points)
    for k, q in enumerate(sample):
        assert_parity(a[q], lit[k], tru[k], mode, "deferred rows %s profile %d" % (mode, q))


@pytest.mark.parametrize("mode", ["O", "X"])
@pytest.mark.parametrize("drift_deg_per_km", [0.01, 0.5])
def test_espace_loop_with_a_rotating_field_angle(mode, drift_deg_per_km):
    """Real IGRF fields turn with height.  0.01 deg/km is the second-order rotation path (FastS), 0.5 deg/km the
    eighth-order one (FastL); both keep the node form (coordinate + slopes) in E-space, unlike the constant-angle path.
    Single profile (segments) and a 200-profile batch (queue kernel) against the oracle."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    alt, freq = synth.default_alt(), synth.default_freq()[::4]
    lat, lon = synth.grid_subset(200)
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    bpsi = bpsi + drift_deg_per_km * (alt - alt[0])[None, :]
    n_points = 6001
    got1 = prhf.vertical_forward_operator(freq, den[3], bmag[3], bpsi[3], alt, mode, n_points)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    got = prhf.vertical_forward_operator_batched(*t, mode, n_points).cpu().numpy()
    sample = [3, 100, 199]
    lit, tru = _truth_and_literal(freq, den[sample], bmag[sample], bpsi[sample], alt, mode, n_points)
    assert_parity(got1, lit[0], tru[0], mode, "rotating angle, single profile %s" % mode)
    for k, q in enumerate(sample):
        assert_parity(got[q], lit[k], tru[k], mode, "rotating angle %g deg/km %s profile %d" % (drift_deg_per_km, mode, q))


@pytest.mark.parametrize("n_prof,n_freq,n_points", [(150, 300, 4100), (200, 1, 5000), (148, 33, 8191), (1000, 7, 4097)])
def test_queued_mode_shapes(n_prof, n_freq, n_points, monkeypatch):
    """Shapes around the edges of the queued decomposition: more frequencies than the 256 threads of a row-setup CTA
    (several CTAs per profile append to the queue), a single frequency, exactly 148 profiles, many short profiles'
    worth of rows; X-mode against the one-CTA-per-row form."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import _cabi, synth
    alt = synth.default_alt()
    freq = np.linspace(0.3, 16.0, n_freq) if n_freq > 1 else np.array([5.0])
    lat, lon = synth.grid_subset(n_prof)
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    dev = torch.device("cuda:0")

    def run():
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
        return prhf.vertical_forward_operator_batched(*t, "X", n_points).cpu().numpy()

    a = run()
    monkeypatch.setenv("PRHF_QUEUE", "0")
    monkeypatch.setattr(_cabi, "_contexts", {})
    b = run()
    assert a.shape == (n_prof, n_freq)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(a)
    assert m.any() and np.max(np.abs(a[m] - b[m]) / np.abs(b[m])) < 1e-12
    lit, tru = _truth_and_literal(freq, den[:2], bmag[:2], bpsi[:2], alt, "X", n_points)
    for q in range(2):
        assert_parity(a[q], lit[q], tru[q], "X", "shape %dx%dx%d profile %d" % (n_prof, n_freq, n_points, q))


@pytest.mark.parametrize("mode,grid", [("X", "uniform"), ("O", "uniform"), ("X", "jittered"), ("O", "jittered")])
def test_every_decomposition_writes_every_row(mode, grid):
    """Shapes on both sides of every switch between the work decompositions (solo <-> planned <-> one CTA per row <->
    queued; row-per-warp <-> tile kernels; warp- <-> thread-per-frequency row setup): the output starts as a sentinel,
    every entry must be overwritten, and sampled rows must equal the single-profile call.  (Found in round 2: 148 ... 222
    profiles of ONE frequency ran the single-launch kernel with the thread-per-frequency flag set and left their
    reflecting rows unwritten.)"""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    alt = synth.default_alt()
    if grid == "jittered":                                    # not uniform: the table-reading m-space loop
        alt = alt + np.random.default_rng(5).uniform(-0.3, 0.3, alt.size)
    dev = torch.device("cuda:0")
    lat, lon = synth.grid_subset(512)
    den_all, bmag_all, bpsi_all = synth.profiles_at(lat, lon, alt)
    t_alt = torch.from_numpy(alt).to(dev)
    checked = 0
    for n_freq in ((1, 2, 9, 174, 257) if grid == "uniform" else (1, 9, 174)):
        freq = np.linspace(0.5, 14.0, n_freq) if n_freq > 1 else np.array([5.0])
        t_freq = torch.from_numpy(freq).to(dev)
        for n_prof in (1, 2, 3, 23, 24, 25, 147, 148, 149, 222, 223, 300, 512):
            if n_prof * n_freq > 60000:
                continue
            t = [torch.from_numpy(np.ascontiguousarray(v[:n_prof])).to(dev) for v in (den_all, bmag_all, bpsi_all)]
            for n_points in (1, 2, 200, 2047, 2048, 4096, 4097, 5000):
                out = torch.full((n_prof, n_freq), -7.0, dtype=torch.float64, device=dev)
                prhf.vertical_forward_operator_batched(t_freq, t[0], t[1], t[2], t_alt, mode, n_points, out=out)
                a = out.cpu().numpy()
                label = "P=%d F=%d n=%d" % (n_prof, n_freq, n_points)
                assert not (a == -7.0).any(), label + ": %d rows never written" % int((a == -7.0).sum())
                if n_points in (200, 2048, 5000):
                    q = n_prof - 1
                    one = prhf.vertical_forward_operator(freq, den_all[q], bmag_all[q], bpsi_all[q], alt, mode, n_points)
                    assert np.array_equal(np.isnan(one), np.isnan(a[q])), label
                    m = np.isfinite(one)
                    assert np.max(np.abs(one[m] - a[q][m]) / np.abs(one[m]), initial=0.0) < (1e-11 if mode == "X" else 1e-9), label
                checked += 1
    assert checked > 150


@pytest.mark.parametrize("mode", ["O", "X"])
def test_queued_mode_with_per_profile_grids_and_frequencies(mode):
    """[P, A] altitude grids and [P, F] frequency sets (every profile its own), 300 profiles at 5000 points per row:
    the queue kernel must index both with the profile's stride.  Against the single-profile call."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    rng = np.random.default_rng(11)
    alt0 = synth.default_alt()
    n_prof, n_freq, n_points = 300, 40, 5000
    lat, lon = synth.grid_subset(n_prof)
    shift = rng.uniform(-3.0, 3.0, n_prof)
    alt = alt0[None, :] + shift[:, None]                      # still uniform per profile, different origin
    den = np.empty((n_prof, alt0.size)); bmag = np.empty_like(den); bpsi = np.empty_like(den)
    for q in range(n_prof):
        d, b, p = synth.profiles_at(lat[q:q + 1], lon[q:q + 1], alt[q])
        den[q], bmag[q], bpsi[q] = d[0], b[0], p[0]
    freq = np.sort(rng.uniform(0.5, 14.0, (n_prof, n_freq)), axis=1)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    out = torch.full((n_prof, n_freq), -7.0, dtype=torch.float64, device=dev)
    prhf.vertical_forward_operator_batched(*t, mode, n_points, out=out)
    a = out.cpu().numpy()
    assert not (a == -7.0).any()
    for q in (0, 1, 149, 150, 299):
        one = prhf.vertical_forward_operator(freq[q], den[q], bmag[q], bpsi[q], alt[q], mode, n_points)
        assert np.array_equal(np.isnan(one), np.isnan(a[q])), q
        m = np.isfinite(one)
        assert np.max(np.abs(one[m] - a[q][m]) / np.abs(one[m]), initial=0.0) < (1e-11 if mode == "X" else 1e-9), q

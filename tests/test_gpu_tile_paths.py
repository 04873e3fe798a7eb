"""Round-2 decompositions of the tile kernel: the E-space grid loop (the stretched grid as a geometric sequence, no table
reads), the live-row queue of large batches, and the rows the row setup finishes itself.

* odd / even n_points, tile boundaries and re-seed boundaries of the E-space loop against the scalar C restatement of
  library.py:459-509 (X-mode) and the long-double truth (O-mode) -- the last pair of points of a row is added outside
  the loop and differs between odd and even n_points;
* queued mode (PRHF_QUEUE=2, default) == one tile-kernel CTA per row (PRHF_QUEUE=0), bit for bit, on a batch that has
  rows without reflection, rows clamped to the first level (finished by the row setup in queued mode) and ordinary rows.
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
    with one tile-kernel CTA per row.  Same bits; oracle on sampled profiles."""
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
    assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(a[7]).all() and np.isnan(a[11]).all()
    if not literal:
        sample = [0, 150, 299]
        lit, tru = _truth_and_literal(freq, den[sample], bmag[sample], bpsi[sample], alt, mode, n_points)
        for k, q in enumerate(sample):
            assert_parity(a[q], lit[k], tru[k], mode, "queued %s profile %d" % (mode, q))

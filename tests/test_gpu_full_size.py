"""BASELINE.json configs[2] and configs[3] at FULL size on the GPU (driver-visible evidence).

The oracle needs ~1 s per profile at n_points = 20000, so the full batches are checked through size-independent
properties -- no status errors, every row of a profile equals the single-profile call bit for bit, the sharded and the
plain batched operator agree bit for bit, finite fraction in the expected band -- and against the oracle / the
long-double truth on sampled profiles.
"""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu


def _device_inputs(lat, lon, params=None):
    import torch
    from pyrayhf_b200 import synth
    alt, freq = synth.default_alt(), synth.default_freq()
    if params is None:
        fof2, hmf2, scale_h, foe = synth.layer_parameters(lat, lon)
        params = np.stack([fof2, hmf2, scale_h, foe, lat], axis=1)
    dev = torch.device("cuda:0")
    den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
    return freq, alt, params, den, bmag, bpsi, torch.from_numpy(freq).to(dev), torch.from_numpy(alt).to(dev)


def _spot_check(vh, dev_profiles, sample, freq, alt, mode, n_points, label):
    """Oracle on the SAME profile arrays the GPU used (the device-built ones, copied back for the sampled profiles: the
    CUDA and the numpy exp differ in the last bits, which may flip a row that sits on its critical frequency)."""
    from oracle import scalar, vfo_oracle
    mult = vfo_oracle.stretch_multiplier(n_points)
    d, b, p = (v[sample].cpu().numpy() for v in dev_profiles)
    truth = scalar.vertical_forward_operator_batched(freq, d, b, p, alt, mode, n_points, variant=1, multiplier=mult)[0]
    lit = scalar.vertical_forward_operator_batched(freq, d, b, p, alt, mode, n_points, variant=0, multiplier=mult)[0]
    for k, q in enumerate(sample):
        # the numpy restatement (bit-identical to the reference) on the first two, the scalar C restatement on the rest
        ref = vfo_oracle.vertical_forward_operator(freq, d[k], b[k], p[k], alt, mode, n_points) if k < 2 else lit[k]
        assert_parity(vh[q], ref, truth[k], mode, "%s profile %d" % (label, q))


@pytest.mark.parametrize("mode", ["X", "O"])
def test_config3_global_grid_at_n20000(mode):
    """65 341 profiles (1-degree global grid) x 174 frequencies, n_points = 20000, device-resident."""
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    lat, lon = synth.global_grid_points()
    freq, alt, params, den, bmag, bpsi, t_freq, t_alt = _device_inputs(lat, lon)
    assert den.shape == (65341, 620)
    vh, st = prhf.vertical_forward_operator_batched(t_freq, den, bmag, bpsi, t_alt, mode, 20000, return_status=True)
    assert int(st.abs().sum().item()) == 0
    a = vh.cpu().numpy()
    assert np.nanmin(a) >= alt[0] and 0.25 < np.isfinite(a).mean() < 0.45
    sample = np.array([0, 32670, 65340, 12345, 40000, 52000])          # poles, equator, mid-latitudes
    for q in sample[:3]:                                               # batched row == single-profile call, bit for bit
        one = prhf.vertical_forward_operator(freq, den[q].cpu().numpy(), bmag[q].cpu().numpy(), bpsi[q].cpu().numpy(),
                                             alt, mode, 20000)
        assert np.array_equal(np.isnan(one), np.isnan(a[q]))
        m = np.isfinite(one)
        assert np.max(np.abs(one[m] - a[q][m]) / np.abs(a[q][m]), initial=0.0) < 1e-12   # other tiling, same rows
    _spot_check(a, (den, bmag, bpsi), sample, freq, alt, mode, 20000, "config3 " + mode)


def test_config4_full_ensemble_member_through_the_sharded_operator():
    """One member of the configs[3] ensemble: 8 192 perturbed profiles, X-mode, n_points = 20000, built on the device from
    40 bytes per profile and run through the sharded operator (one rank), result in the page-locked host buffer."""
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import sharding, synth
    lat, lon = synth.grid_subset(8192)
    params = np.stack(synth.ensemble_member_parameters(lat, lon, 17), axis=1)
    alt, freq = synth.default_alt(), synth.default_freq()
    op = sharding.ShardedForwardOperator(8192, freq.size, layout="interleaved", gather_to=0)
    vh = np.array(op.from_parameters(freq, params, alt, "X", 20000), copy=True)
    op.close()
    assert vh.shape == (8192, 174) and 0.2 < np.isfinite(vh).mean() < 0.5
    # the plain batched operator on the same device-built profiles gives the same bits
    dev = torch.device("cuda:0")
    den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
    ref = prhf.vertical_forward_operator_batched(torch.from_numpy(freq).to(dev), den, bmag, bpsi,
                                                 torch.from_numpy(alt).to(dev), "X", 20000).cpu().numpy()
    assert np.array_equal(vh, ref, equal_nan=True)
    _spot_check(vh, (den, bmag, bpsi), np.array([0, 4095, 8191, 1234]), freq, alt, "X", 20000, "config4 member 17")

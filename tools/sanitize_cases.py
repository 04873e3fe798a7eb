#!/usr/bin/env python
"""Small invocations of every kernel family, meant to be run under compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py

Sizes are kept small (the tools slow kernels down 10-100x); every decomposition of the operator is hit:
single-launch kernel, planned mode with and without programmatic dependent launch, direct tile kernel,
row-per-warp kernel, lane-mode row setup, the host entry with CUDA-graph replay, the streaming entry and the
stage / tracer / residual operators.  Results are compared with the numpy port so that a silent corruption shows.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import pyrayhf_b200 as prhf
    from pyrayhf_b200 import synth
    from oracle import vfo_oracle

    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    alt = synth.default_alt()
    freq = synth.default_freq()
    fsub = np.ascontiguousarray(freq[::6])

    def check(tag, got, ref, tol=1e-9, mode="X"):
        got, ref = np.asarray(got), np.asarray(ref)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), tag
        m = np.isfinite(ref)
        err = float(np.max(np.abs(got[m] - ref[m]) / np.abs(ref[m]))) if m.any() else 0.0
        assert err <= tol, (tag, err)
        print("ok  %-44s max rel err %.2e" % (tag, err), flush=True)

    den, bmag, bpsi, _ = synth.bench_day_profile()
    # 1. single-launch kernel through the numpy drop-in (3 calls: plain, capture, graph replay)
    ref = vfo_oracle.vertical_forward_operator(fsub, den, bmag, bpsi, alt, "X", 4096)
    for k in range(3):
        got = prhf.vertical_forward_operator(fsub, den, bmag, bpsi, alt, "X", 4096)
    check("solo X n=4096 (host entry, graph replay)", got, ref)
    got = prhf.vertical_forward_operator(fsub, den, bmag, bpsi, alt, "O", 4096)
    ref_o = vfo_oracle.vertical_forward_operator(fsub, den, bmag, bpsi, alt, "O", 4096)
    check("solo O n=4096", got, ref_o, tol=1e-4)

    def batch(n_prof, n_points, mode="X", fr=fsub, tol=1e-9):
        lat, lon = synth.grid_subset(n_prof)
        d2, b2, p2 = synth.profiles_at(lat, lon, alt)
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (fr, d2, b2, p2, alt)]
        vh = prhf.vertical_forward_operator_batched(*t, mode, n_points, errors="nan").cpu().numpy()
        k = min(n_prof, 3)
        ref = vfo_oracle.vertical_forward_operator_batched(fr, d2[:k], b2[:k], p2[:k], alt, mode, n_points)
        check("batched P=%d n=%d %s" % (n_prof, n_points, mode), vh[:k], ref, tol=tol)
        return d2, b2, p2

    batch(2, 4096)          # planned mode, programmatic dependent launch
    batch(6, 2048)          # planned mode, plain stream order
    batch(32, 4100)         # direct mode, tile kernel
    batch(32, 200)          # direct mode, row-per-warp kernel
    d2, b2, p2 = batch(160, 64, fr=np.ascontiguousarray(freq[::3]))   # lane-mode row setup
    batch(3, 2048, mode="O", tol=1e-4)

    # host entry for a batch (packed arena) and, when present, the streaming entry
    vh = prhf.vertical_forward_operator_batched(fsub, d2[:5], b2[:5], p2[:5], alt, "X", 300, errors="nan")
    ref = vfo_oracle.vertical_forward_operator_batched(fsub, d2[:5], b2[:5], p2[:5], alt, "X", 300)
    check("host batched P=5 n=300", vh, ref)
    if hasattr(prhf, "vertical_forward_operator_streamed"):
        vh = prhf.vertical_forward_operator_streamed(fsub, d2[:37], b2[:37], p2[:37], alt, "X", 300, chunk_profiles=8)
        ref = vfo_oracle.vertical_forward_operator_batched(fsub, d2[:37], b2[:37], p2[:37], alt, "X", 300)
        check("streamed P=37 n=300 (chunks of 8)", vh, ref)

    # stages, tracers, residual
    rg = prhf.regrid_to_nonuniform_grid(fsub[3:9] * 1e6, den, bmag, bpsi, alt, mode="X", n_points=64)
    ro = vfo_oracle.regrid_dict(fsub[3:9] * 1e6, den, bmag, bpsi, alt, "X", 64)
    for key in ("alt", "den", "bmag", "bpsi", "dist", "crit_height"):
        assert np.allclose(rg[key], ro[key], rtol=1e-12, atol=1e-12, equal_nan=True), key
    print("ok  regrid stage", flush=True)
    rays = prhf.trace_rays_snells_batched(np.array([3e6, 4e6, 9e6]), np.array([30.0, 60.0, 45.0]), alt, den, bmag,
                                          bpsi, "O", geometry="spherical")
    assert np.isfinite(rays["group_path_km"]).any()
    rays = prhf.trace_rays_snells_batched(np.array([3e6, 4e6, 9e6]), np.array([30.0, 60.0, 45.0]), alt, den, bmag,
                                          bpsi, "X", geometry="cartesian")
    print("ok  Snell tracers", flush=True)
    vm = np.random.default_rng(1).random((9, fsub.size)) * 300 + 100
    vm[2, 3] = np.nan
    res, chi2 = prhf.residual_VH_batched(vm[0], vm)
    assert np.isfinite(chi2).all()
    print("ok  residual", flush=True)
    X = np.linspace(0.0, 0.9, 100)
    mu, mup = prhf.find_mu_mup(X, np.full_like(X, 0.2), np.full_like(X, 30.0), "O")
    assert np.isfinite(mu).all()
    vh1 = prhf.find_vh(np.tile(X, (4, 1)), np.full((4, 100), 0.2), np.full((4, 100), 30.0), np.full((4, 100), 0.5), 80.0,
                       "X")
    assert vh1.shape == (4,)
    print("ok  find_mu_mup / find_vh", flush=True)
    if hasattr(prhf, "brute_force_search"):
        pass
    torch.cuda.synchronize()
    print("ALL CASES OK", flush=True)


if __name__ == "__main__":
    main()

"""Developer probe (not part of the product): kernel timings for a few shapes + FP64 peak."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200  # noqa: E402
from pyrayhf_b200 import _cabi, synth  # noqa: E402

_vp = ctypes.c_void_p


def time_device(P, mode, n, reps=10, literal=False):
    dev = torch.device("cuda:0")
    alt = synth.default_alt()
    freq = synth.default_freq()
    if P == 1:
        den, bmag, bpsi, _ = synth.single_day_profile()
        den, bmag, bpsi = den[None], bmag[None], bpsi[None]
    else:
        lat, lon = synth.grid_subset(P)
        den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
    out = torch.empty((P, freq.size), dtype=torch.float64, device=dev)
    for _ in range(3):
        pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=out, literal=literal, errors='nan')
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=out, literal=literal, errors='nan')
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = np.array(ts)
    vh = out.cpu().numpy()
    live = int(np.isfinite(vh).sum())
    return dict(P=P, mode=mode, n=n, literal=literal, ms_min=float(ts.min()), ms_med=float(np.median(ts)),
                vh_per_s=P * freq.size / (np.median(ts) * 1e-3), live_rows=live,
                gpts_per_s=live * n / (np.median(ts) * 1e-3) / 1e9)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        P, mode, n = int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
        reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
        print(json.dumps(time_device(P, mode, n, reps=reps)))
        return
    ctx = _cabi.context(0)
    print(json.dumps({"fp64_peak_tflops": ctx.measure_fp64_peak()}))
    for args in [(1, 'X', 20000), (1, 'O', 20000), (1, 'X', 200), (64, 'X', 20000), (512, 'X', 20000),
                 (4096, 'X', 200), (4096, 'O', 200), (1, 'X', 20000, 10, True), (64, 'X', 20000, 5, True)]:
        print(json.dumps(time_device(*args)), flush=True)
    # e2e single profile through the numpy drop-in
    den, bmag, bpsi, alt = synth.single_day_profile()
    freq = synth.default_freq()
    for _ in range(5):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, 'X', 20000)
    t0 = time.perf_counter()
    K = 50
    for _ in range(K):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, 'X', 20000)
    dt = (time.perf_counter() - t0) / K
    print(json.dumps({"e2e_single_profile_ms": dt * 1e3, "vh_per_s": freq.size / dt}))


if __name__ == "__main__":
    main()

"""Developer probe: BASELINE config 4 on one GPU -- ensemble members of 8192 perturbed profiles, X-mode, n_points = 20000,
profiles built on the device (prhf_synth_profiles_f64), timed with CUDA events."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n_base = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    lat, lon = synth.grid_subset(n_base)
    freq = torch.from_numpy(synth.default_freq()).to(dev)
    alt = torch.from_numpy(synth.default_alt()).to(dev)
    out = torch.empty((n_base, freq.numel()), dtype=torch.float64, device=dev)
    times, live = [], []
    for member in range(4):
        par = synth.ensemble_member_parameters(lat, lon, member)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        den, bmag, bpsi = synth.profiles_from_parameters_device(*par)
        pyrayhf_b200.vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, 'X', 20000, out=out, errors='nan')
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        live.append(int(torch.isfinite(out).sum().item()))
    ms = float(np.median(times[1:]))
    print(json.dumps({"config": "4 (one GPU's share of one member)", "profiles": n_base, "n_points": 20000, "mode": "X",
                      "ms_per_member_chunk": ms, "vh_per_s": n_base * freq.numel() / (ms * 1e-3),
                      "grid_points_per_s": live[-1] * 20000 / (ms * 1e-3), "finite_rows": live[-1],
                      "h2d_bytes_per_profile": 40}))


if __name__ == "__main__":
    main()

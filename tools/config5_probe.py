"""Developer probe: BASELINE config 5 -- one profile, 1 740 frequencies (0.01 MHz step), n_points 200 ... 50 000, both
modes; device-resident inputs, CUDA events."""
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
    den, bmag, bpsi, alt = synth.bench_day_profile()
    freq = np.arange(0.01, 17.41, 0.01)
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den[None], bmag[None], bpsi[None], alt)]
    out = torch.empty((1, freq.size), dtype=torch.float64, device=dev)
    for mode in "XO":
        for n in (200, 2000, 20000, 50000):
            for _ in range(3):
                pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=out)
            torch.cuda.synchronize()
            ev = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=out)
                b.record()
                ev.append((a, b))
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
            live = int(torch.isfinite(out).sum().item())
            print(json.dumps({"mode": mode, "n_points": n, "n_freq": int(freq.size), "ms": ms,
                              "vh_per_s": freq.size / (ms * 1e-3), "grid_points_per_s": live * n / (ms * 1e-3)}))


if __name__ == "__main__":
    main()

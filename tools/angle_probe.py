"""Developer probe (not part of the product): cost of a field angle that turns with height (real IGRF fields) against
the constant angle of the synthetic dipole.  4096 profiles x 174 freqs, n_points = 20000, device-resident."""
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
    alt, freq = synth.default_alt(), synth.default_freq()
    lat, lon = synth.grid_subset(4096)
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for drift in (0.0, 0.01, 0.5, 5.0):
        psi = bpsi + drift * (alt - alt[0])[None, :]
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, psi, alt)]
        out = torch.empty((4096, freq.size), dtype=torch.float64, device=dev)
        for mode in ("X", "O"):
            for _ in range(2):
                pyrayhf_b200.vertical_forward_operator_batched(*t, mode, 20000, out=out, errors="nan")
            ts = []
            for _ in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pyrayhf_b200.vertical_forward_operator_batched(*t, mode, 20000, out=out, errors="nan")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("angle drift %5.2f deg/km  %s-mode  %.2f ms  (%.1f M vh/s)" % (
                drift, mode, np.median(ts), 4096 * freq.size / np.median(ts) / 1e3))


if __name__ == "__main__":
    main()

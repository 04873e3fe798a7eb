"""Developer probe of the mixed-precision EXPERIMENT (commit 093e345 "experiment: mixed precision mode"; the mode is
not in the product, see profiles/mixed_f32_ab_r02.txt): one ensemble member (8 192 profiles), time and deviation from
the double-precision result, per mode.  Runs only against a library built from that commit."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200 as prhf  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    alt, freq = synth.default_alt(), synth.default_freq()
    lat, lon = synth.grid_subset(8192)
    params = np.stack(synth.ensemble_member_parameters(lat, lon, 0), axis=1)
    den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
    tf, ta = torch.from_numpy(freq).to(dev), torch.from_numpy(alt).to(dev)
    out = torch.empty((8192, freq.size), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            ev.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev]))

    for mode in ("X", "O"):
        for n in (20000, 5000):
            res = {}
            for prec in ("float64", "mixed"):
                ms = timed(lambda: prhf.vertical_forward_operator_batched(tf, den, bmag, bpsi, ta, mode, n, out=out,
                                                                          errors='nan', precision=prec))
                res[prec] = (ms, out.cpu().numpy().copy())
            a, b = res["mixed"][1], res["float64"][1]
            assert np.array_equal(np.isnan(a), np.isnan(b))
            m = np.isfinite(b)
            err = np.abs(a[m] - b[m]) / np.abs(b[m])
            print("%s n=%5d  float64 %.2f ms  mixed %.2f ms  speed-up %.2fx | deviation max %.2e  99.9%% %.2e  median %.2e  "
                  "(%d finite rows, masks identical)" % (mode, n, res["float64"][0], res["mixed"][0],
                                                         res["float64"][0] / res["mixed"][0], err.max(),
                                                         np.quantile(err, 0.999), np.median(err), int(m.sum())), flush=True)


if __name__ == "__main__":
    main()

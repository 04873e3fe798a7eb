"""Developer probe of the streaming entry: device-resident vs page-locked host inputs, chunk sizes, copy rate."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200 as prhf  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402


def main():
    n_prof = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    dev = torch.device("cuda:0")
    alt, freq = synth.default_alt(), synth.default_freq()
    lat, lon = synth.grid_subset(8192)
    params = np.concatenate([np.stack(synth.ensemble_member_parameters(lat, lon, m), axis=1)
                             for m in range((n_prof + 8191) // 8192)])[:n_prof]
    den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
    t_freq, t_alt = torch.from_numpy(freq).to(dev), torch.from_numpy(alt).to(dev)
    h = [prhf.pinned_empty((n_prof, alt.size)) for _ in range(3)]
    for a, d in zip(h, (den, bmag, bpsi)):
        torch.from_numpy(a).copy_(d)
    out = prhf.pinned_empty((n_prof, freq.size))
    torch.cuda.synchronize()
    # raw copy rates
    buf = torch.empty_like(den)
    for _ in range(2):
        buf.copy_(torch.from_numpy(h[0]), non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        buf.copy_(torch.from_numpy(h[0]), non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print("H2D pinned: %.1f GB/s (%.1f MB in %.2f ms)" % (h[0].nbytes / dt / 1e9, h[0].nbytes / 1e6, dt * 1e3))

    def run(label, args, n_points, chunk, reps=3):
        for _ in range(2):
            prhf.vertical_forward_operator_streamed(*args, "X", n_points, errors="nan", out=out, chunk_profiles=chunk)
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            prhf.vertical_forward_operator_streamed(*args, "X", n_points, errors="nan", out=out, chunk_profiles=chunk)
            ts.append(time.perf_counter() - t0)
        ms = 1e3 * float(np.median(ts))
        print("%-34s n=%5d chunk=%5d : %8.2f ms  (%.1f M vh/s)" % (label, n_points, chunk, ms,
                                                                  n_prof * freq.size / ms / 1e3), flush=True)
        return ms

    dev_args = (t_freq, den, bmag, bpsi, t_alt)
    host_args = (freq, h[0], h[1], h[2], alt)
    if len(sys.argv) > 2 and sys.argv[2] == "quick":
        for n_points in (20000, 200):
            run("device inputs -> pinned out", dev_args, n_points, 2048, reps=1)
            run("pinned inputs -> pinned out", host_args, n_points, 2048, reps=1)
        return
    for n_points in (20000, 200):
        for chunk in (0, 1024, 2048, 4096, 8192):
            run("device inputs -> pinned out", dev_args, n_points, chunk)
            run("pinned inputs -> pinned out", host_args, n_points, chunk)
    pag = [np.array(a) for a in h]
    run("pageable inputs -> pinned out", (freq, pag[0], pag[1], pag[2], alt), 20000, 0)
    t0 = time.perf_counter()
    prhf.vertical_forward_operator_batched(freq, pag[0], pag[1], pag[2], alt, "X", 20000, errors="nan")
    print("packed host entry (prhf_vfo_host_f64), pageable inputs: %.2f ms" % (1e3 * (time.perf_counter() - t0)))


if __name__ == "__main__":
    main()

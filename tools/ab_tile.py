"""Developer A/B of library builds on one box (not part of the product).

usage: python tools/ab_tile.py libA.so libB.so:PRHF_NO_QUEUE=1 ...     (paths relative to pyrayhf_b200/csrc)
Every build runs the same cases in its own process (PRHF_LIB_PATH); timings are medians of CUDA-event timed calls with
an L2 flush in between; the virtual heights of every case are compared with the first build's (NaN masks, max relative
difference) and the first case is also compared with the scalar C oracle on two profiles.
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(4096, "X", 20000, 5), (4096, "O", 20000, 5), (1, "X", 20000, 30), (1, "O", 20000, 30), (8, "X", 20000, 10),
         (4096, "X", 200, 10), (64, "X", 5000, 10)]


def child(tag):
    import torch
    sys.path.insert(0, ROOT)
    import pyrayhf_b200
    from pyrayhf_b200 import synth
    dev = torch.device("cuda:0")
    alt, freq = synth.default_alt(), synth.default_freq()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for P, mode, n, reps in CASES:
        if P == 1:
            den, bmag, bpsi, _ = synth.single_day_profile()
            den, bmag, bpsi = den[None], bmag[None], bpsi[None]
        else:
            lat, lon = synth.grid_subset(P)
            den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
        t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den, bmag, bpsi, alt)]
        vh = torch.empty((P, freq.size), dtype=torch.float64, device=dev)
        for _ in range(3):
            pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=vh, errors="nan")
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pyrayhf_b200.vertical_forward_operator_batched(*t, mode, n, out=vh, errors="nan")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        key = "%d_%s_%d" % (P, mode, n)
        out[key] = vh.cpu().numpy()
        print("%-28s %-14s ms med %.4f min %.4f" % (tag, key, float(np.median(ts)), float(np.min(ts))), flush=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tag), **out)


def main():
    libs = sys.argv[1:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tags = []
    for spec in libs:                                  # "lib.so" or "lib.so:ENV=VALUE[:ENV=VALUE]"
        lib, *sets = spec.split(":")
        tag = (os.path.basename(lib).replace("libpyrayhf_b200", "").replace(".so", "") or "head") + "".join(
            "+" + kv.split("=")[0].replace("PRHF_", "").lower() for kv in sets)
        tags.append(tag)
        env = dict(os.environ, PRHF_LIB_PATH=os.path.join(ROOT, "pyrayhf_b200", "csrc", lib))
        env.update(dict(kv.split("=", 1) for kv in sets))
        subprocess.run([sys.executable, __file__, "--child", tag], env=env, check=False, timeout=600)
    base = np.load(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tags[0]))
    for tag in tags[1:]:
        f = os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tag)
        if not os.path.exists(f):
            print(tag, "produced no output")
            continue
        d = np.load(f)
        for k in base.files:
            a, b = base[k], d[k]
            same_mask = bool(np.array_equal(np.isnan(a), np.isnan(b)))
            m = np.isfinite(a) & np.isfinite(b)
            rel = float(np.max(np.abs(a[m] - b[m]) / np.abs(a[m]), initial=0.0))
            print("%-20s vs %-8s %-14s masks %s  max rel diff %.3e" % (tag, tags[0], k, "same" if same_mask else "DIFFER", rel))
    # truth check of the last build on the single-profile cases
    sys.path.insert(0, ROOT)
    from oracle import scalar, vfo_oracle
    from pyrayhf_b200 import synth
    den, bmag, bpsi, alt = synth.single_day_profile()
    freq = synth.default_freq()
    mult = vfo_oracle.stretch_multiplier(20000)
    for tag in tags:
        d = np.load(os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % tag))
        for mode in "XO":
            tru = scalar.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, 20000, variant=1, multiplier=mult)
            got = d["1_%s_20000" % mode][0]
            m = np.isfinite(tru)
            print("%-20s single %s vs long-double truth: masks %s max rel %.3e" % (
                tag, mode, np.array_equal(np.isnan(got), np.isnan(tru)), float(np.max(np.abs(got[m] - tru[m]) / np.abs(tru[m])))))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()

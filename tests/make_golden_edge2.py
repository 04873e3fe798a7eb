"""Generates tests/golden/edge2.npz from the LIVE reference (dev container only): contract edges added in round 2.

* altitude grids that are NOT increasing: the reference does not reject them; np.interp's range tests
  (binary_search_with_guess) decide, so a strictly decreasing grid gives constant interpolants and mostly-NaN rows;
* profiles with more levels than the GPU path stages in shared memory (3 000 and 5 000 levels): the reference has no
  limit (np.interp, library.py:424-426).

    python tests/make_golden_edge2.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = ref_import.load_reference_library()
    out = {}
    freq = synth.default_freq()
    den, bmag, bpsi, alt = synth.bench_day_profile()
    rev = [np.ascontiguousarray(v[::-1]) for v in (den, bmag, bpsi, alt)]
    for mode in ("O", "X"):
        for n in (200, 2000):
            out["reversed_%s_%d" % (mode, n)] = ref.vertical_forward_operator(freq, rev[0], rev[1], rev[2], rev[3], mode, n)
            out["altdown_%s_%d" % (mode, n)] = ref.vertical_forward_operator(freq, den, bmag, bpsi, rev[3], mode, n)
    # a night profile reversed as well (different peak position)
    dn, bn, pn = synth.profiles_at([35.0], [170.0], alt)
    out["night_den"], out["night_bmag"], out["night_bpsi"] = dn[0], bn[0], pn[0]
    for mode in ("O", "X"):
        out["night_reversed_%s_200" % mode] = ref.vertical_forward_operator(
            freq, dn[0][::-1].copy(), bn[0][::-1].copy(), pn[0][::-1].copy(), rev[3], mode, 200)
    # long profiles
    fsub = np.ascontiguousarray(freq[::5])
    out["fsub"] = fsub
    for n_alt in (3000, 5000):
        a = np.linspace(80.0, 700.0, n_alt)
        d, b, p = synth.profiles_at([4.5, -40.0], [0.0, 120.0], a)
        p = p + np.linspace(0.0, 3.0, n_alt)[None, :] * np.array([[0.0], [1.0]])   # second profile: rotating field angle
        out["long%d_alt" % n_alt], out["long%d_den" % n_alt] = a, d
        out["long%d_bmag" % n_alt], out["long%d_bpsi" % n_alt] = b, p
        for mode in ("O", "X"):
            for n in (200, 5000):
                out["long%d_%s_%d" % (n_alt, mode, n)] = np.stack(
                    [ref.vertical_forward_operator(fsub, d[q], b[q], p[q], a, mode, n) for q in range(2)])
    for k in sorted(out):
        if k.startswith(("reversed", "altdown", "night_rev")):
            print(k, int(np.isfinite(out[k]).sum()), "finite of", out[k].size)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "edge2.npz"), **out)


if __name__ == "__main__":
    main()

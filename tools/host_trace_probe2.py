"""Developer probe: what makes cudaGraphLaunch cheap (3.8 us) or expensive (13.8 us) for back-to-back host-entry calls."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200  # noqa: E402
from pyrayhf_b200 import synth  # noqa: E402

den, bmag, bpsi, alt = synth.bench_day_profile()
freq = synth.default_freq()
tiny = torch.zeros(16, device="cuda:0")


def spin(us):
    t = time.perf_counter()
    while (time.perf_counter() - t) * 1e6 < us:
        pass


variants = {
    "nothing": lambda: None,
    "spin 100 us": lambda: spin(100),
    "spin 1000 us": lambda: spin(1000),
    "torch.cuda.synchronize": lambda: torch.cuda.synchronize(),
    "tiny kernel + synchronize": lambda: (tiny.add_(1), torch.cuda.synchronize()),
    "tiny kernel, no synchronize": lambda: tiny.add_(1),
}
for name, between in variants.items():
    for _ in range(10):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, "X", 20000)
    tot = 0.0
    for _ in range(128):
        between()
        t0 = time.perf_counter()
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, "X", 20000)
        tot += time.perf_counter() - t0
    print("between calls: %-30s python-level mean %.2f us per call" % (name, 1e6 * tot / 128), flush=True)

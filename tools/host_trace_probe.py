"""Developer probe: PRHF_HOST_TRACE=1 python tools/host_trace_probe.py -- where a single-profile call through the
numpy drop-in spends its wall-clock time (packing, enqueue = graph launch, wait = stream sync, unpacking), with and
without an L2 flush between calls."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for with_flush in (False, True):
    for _ in range(10):
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, "X", 20000)
    tot = 0.0
    for _ in range(128):
        if with_flush:
            flush.zero_()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, "X", 20000)
        tot += time.perf_counter() - t0
    print("L2 flush between calls: %s   python-level mean %.2f us per call" % (with_flush, 1e6 * tot / 128), flush=True)

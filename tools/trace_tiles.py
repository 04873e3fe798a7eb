"""Developer tool: per-CTA phase timeline of the tile kernel for the bench workload (needs `make trace`)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyrayhf_b200 import _cabi, synth  # noqa: E402

_cabi.LIB_PATH = os.path.join(ROOT, "pyrayhf_b200", "csrc", "libpyrayhf_b200_trace.so")
import pyrayhf_b200  # noqa: E402

vp = ctypes.c_void_p


def main():
    dev = torch.device("cuda:0")
    den, bmag, bpsi, alt = synth.bench_day_profile()
    freq = synth.default_freq()
    t = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (freq, den[None], bmag[None], bpsi[None], alt)]
    ctx = _cabi.context(0)
    L = ctx.lib
    n_tiles = 174 * 32
    L.prhf_debug_trace_alloc.argtypes = [vp, ctypes.c_int64]
    L.prhf_debug_trace_read.argtypes = [vp, ctypes.c_int64, vp]
    for _ in range(3):
        pyrayhf_b200.vertical_forward_operator_batched(*t, 'X', 20000)
    torch.cuda.synchronize()
    L.prhf_debug_trace_alloc(ctx.handle, n_tiles)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if not os.environ.get('TRACE_NO_FLUSH'):
        flush.zero_()
    pyrayhf_b200.vertical_forward_operator_batched(*t, 'X', 20000)
    out = np.zeros((n_tiles + 4096, 8), dtype=np.int64)
    L.prhf_debug_trace_read(ctx.handle, n_tiles, vp(out.ctypes.data))
    k1 = out[n_tiles:]
    extra = k1.reshape(-1)[16384:16384 + 16 * 1024].reshape(-1, 16)
    k1 = k1[:1024]
    extra = extra[:len(k1)][k1[:, 0] != 0]
    k1 = k1[k1[:, 0] != 0]
    out = out[:n_tiles]
    print('K1 CTAs', len(k1))
    k1_t0 = k1[:, 6].min()
    print('  K1 start spread %.2f us, K1 last end at %.2f us after first start' % ((k1[:, 6].max() - k1_t0) / 1e3, (k1[:, 7].max() - k1_t0) / 1e3))
    for a, b, nm in ((0, 1, 'stage loads'), (1, 2, 'argmax/min'), (2, 3, 'checks/flags'), (3, 4, 'rows')):
        d = k1[:, b] - k1[:, a]
        print('  K1 %-12s cycles: median %6.0f max %6.0f' % (nm, np.median(d), d.max()))
    print('  K1 CTA total median %d max %d' % (np.median(k1[:, 4] - k1[:, 0]), (k1[:, 4] - k1[:, 0]).max()))
    def seg(name, a, b):
        ok = (a != 0) & (b != 0)
        if ok.any():
            print('    %-34s median %6.0f  (n=%d)' % (name, np.median((b - a)[ok]), ok.sum()))
    seg('argmax end -> checks loop end', k1[:, 2], extra[:, 0])
    seg('checks reduction', extra[:, 0], extra[:, 1])
    seg('record + sync', extra[:, 1], k1[:, 3])
    seg('solo: kx/ky + screen loop', k1[:, 3], extra[:, 3])
    seg('solo: first reduction', extra[:, 3], extra[:, 4])
    seg('solo: literal eval + deep loop', extra[:, 4], extra[:, 5])
    seg('solo: second reduction shuffles', extra[:, 5], extra[:, 9])
    seg('solo: second reduction barrier+', extra[:, 9], extra[:, 6])
    seg('solo: -> row scales (warp loop)', extra[:, 6], extra[:, 7])
    seg('solo: -> h_c', extra[:, 7], extra[:, 8])
    seg('solo: stores', extra[:, 8], k1[:, 4])
    seg('tile: K1 end -> row constants', k1[:, 4], extra[:, 10])
    used = out[:, 1] != 0
    tr = out[used]
    live = tr[:, 7] != 0
    print("tiles launched", used.sum(), "live", live.sum())
    g0 = tr[:, 1].min()
    gt_end = tr[:, 0] >> 10
    tr[:, 0] &= 1023
    print("tile starts: first %.2f us, last %.2f us after K1 first start; last tile end %.2f us" % (
        (tr[:, 1].min() - k1_t0) / 1e3, (tr[:, 1].max() - k1_t0) / 1e3, (gt_end[gt_end > 0].max() - k1_t0) / 1e3))
    print("globaltimer span of CTA starts: %.1f us" % ((tr[:, 1].max() - g0) / 1e3))
    lt = tr[live]
    names = ["span-load", "window", "staging", "loop", "reduce"]
    for k, nm in enumerate(names):
        d = lt[:, 3 + k] - lt[:, 2 + k]
        print("%-10s cycles: median %7.0f  p10 %7.0f  p90 %7.0f  max %7.0f" % (nm, np.median(d), np.percentile(d, 10),
                                                                          np.percentile(d, 90), d.max()))
    tot = lt[:, 7] - lt[:, 2]
    print("tile total cycles: median %.0f p90 %.0f max %.0f" % (np.median(tot), np.percentile(tot, 90), tot.max()))
    dead = tr[~live]
    # start-time histogram in 2 us bins
    st = (tr[:, 1] - g0) / 1e3
    hist, edges = np.histogram(st, bins=np.arange(0, st.max() + 2, 2.0))
    print("CTA starts per 2us bin:", hist.tolist())
    # per-SM busy: number of tiles per SM
    sm = lt[:, 0]
    cnt = np.bincount(sm, minlength=148)
    print("live tiles per SM: min %d max %d" % (cnt.min(), cnt.max()))


if __name__ == "__main__":
    main()

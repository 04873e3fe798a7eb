"""Developer repro script."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyrayhf_b200 import _cabi
_cabi.LIB_PATH = os.path.join(ROOT, "pyrayhf_b200", "csrc", "libpyrayhf_b200_trace.so")
import pyrayhf_b200
from pyrayhf_b200 import synth
import ctypes
def dump(n_rows, max_seg=19):
    ctx = _cabi.context(-1)
    rs = np.zeros(n_rows); cnt = np.zeros(n_rows, dtype=np.uint32); plan = np.zeros(4, dtype=np.int32)
    tiles = np.zeros((n_rows * max_seg, 2), dtype=np.int32)
    vp = ctypes.c_void_p
    ctx.lib.prhf_debug_dump.argtypes = [vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int64]
    ctx.lib.prhf_debug_dump(ctx.handle, n_rows, vp(rs.ctypes.data), vp(cnt.ctypes.data), vp(plan.ctypes.data), vp(tiles.ctypes.data), n_rows * max_seg)
    live = np.isfinite(rs)
    print("  plan", plan.tolist(), "live", live.sum(), "nonzero counters", np.flatnonzero(cnt)[:10].tolist(), cnt[np.flatnonzero(cnt)[:10]].tolist(),
          "span[1740],[1914]", rs[1740], rs[1914])
    nt = plan[0]
    rows_in_list = np.unique(tiles[:nt, 0])
    print("  rows in list", rows_in_list.size, "1740 in list", 1740 in rows_in_list, "1914 in list", 1914 in rows_in_list, flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "literal"
alt = synth.default_alt(); f = synth.default_freq()
g = np.load(os.path.join(ROOT, "tests/golden/synthetic.npz"))
den, bmag, bpsi = synth.profiles_at(g["lat"], g["lon"], g["alt"])
sub = g["sub_20000"]
def planned_check(tag):
    got = pyrayhf_b200.vertical_forward_operator_batched(g["freq"], den[sub], bmag[sub], bpsi[sub], g["alt"], 'O', 20000)
    ref = g["ref_O_20000"]
    bad = np.argwhere(np.isnan(got) != np.isnan(ref))
    print(tag, "mask mismatches", len(bad), bad[:6].tolist(), "got", got[10:12, 0], "ref", ref[10:12, 0], flush=True)
    dump(16 * 174)
if which == "literal":
    d1, b1, p1, _ = synth.single_day_profile()
    vh = pyrayhf_b200.vertical_forward_operator(f, d1, b1, p1, alt, 'X', 20000, literal=True)
    print("literal alone", np.isfinite(vh).sum(), flush=True)
elif which == "seq2":
    planned_check("fresh")
    planned_check("second")
    planned_check("third")
elif which == "seq":
    planned_check("fresh")
    got = pyrayhf_b200.vertical_forward_operator_batched(g["freq"], den, bmag, bpsi, g["alt"], 'O', 200)
    print("direct n=200", np.isfinite(got).sum(), flush=True)
    dump(16 * 174)
    planned_check("after direct")
    planned_check("again")

"""Developer tool: extra seeds of tests/test_gpu_fuzz.py::test_random_profiles (n_points <= 4097 keeps the CPU oracle fast)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import pyrayhf_b200  # noqa: E402
import test_gpu_fuzz as T  # noqa: E402
from conftest import assert_parity  # noqa: E402
from oracle import scalar, vfo_oracle  # noqa: E402

first, count = int(sys.argv[1]), int(sys.argv[2])
bad = n_cases = 0
worst = {'O': 0.0, 'X': 0.0}
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    for kind in range(40):
        freq, den, bmag, bpsi, alt = T.random_profile(rng, kind + 7 * seed)
        n = int(rng.choice([1, 2, 3, 7, 50, 333, 2049, 4097]))
        for mode in "OX":
            m = vfo_oracle.stretch_multiplier(n)
            try:
                lit = scalar.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n, variant=0, multiplier=m)
            except (ValueError, IndexError):
                continue
            tru = scalar.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n, variant=1, multiplier=m)
            got = pyrayhf_b200.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n)
            n_cases += 1
            try:
                assert_parity(got, lit, tru, mode, label="seed %d kind %d n %d %s" % (seed, kind, n, mode))
            except AssertionError as exc:
                bad += 1
                print(str(exc)[:200])
            ok = np.isfinite(lit)
            want = lit if mode == 'X' else tru
            if ok.any():
                worst[mode] = max(worst[mode], float(np.nanmax(np.abs(got[ok] - want[ok]) / np.abs(want[ok]))))
print("extra fuzz: %d cases, %d failures, worst rel err %s" % (n_cases, bad, worst))

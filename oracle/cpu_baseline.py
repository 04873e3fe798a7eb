"""TEST/BENCH INFRASTRUCTURE ONLY -- timing harness for the CPU legs of bench.py.

What is timed is the reference's own ``PyRayHF.library.vertical_forward_operator`` (library.py:459-509), loaded
from the unmodified copy in ``oracle/_ref`` (``oracle/make_ref.py``; ``kind = "reference"``).  Only when that copy
is absent does it fall back to the numpy restatement ``oracle/vfo_oracle.py`` (``kind = "port"``: the same
whole-array numpy evaluation, bit-identical results in the dev container, ~25 % faster because it evaluates sin/cos
once).  Two figures, as BASELINE.md asks: (a) ONE process, how the reference actually runs; (b) one process per
host core, one profile per task, the best the reference can do without code changes ("all the host threads it can
use": the reference is single-threaded numpy).  Workers are spawned (never forked from a CUDA process) and import
numpy + the reference/oracle only -- nothing of the product package.
"""
import multiprocessing as mp
import os
import time
import warnings

import numpy as np

_FN = None


def _operator():
    """(callable, kind) -- resolved once per process."""
    global _FN
    if _FN is None:
        from oracle import make_ref
        ref = make_ref.load()
        if ref is not None:
            _FN = (ref.vertical_forward_operator, "reference")
        else:
            from oracle import vfo_oracle
            _FN = (vfo_oracle.vertical_forward_operator, "port")
    return _FN


def baseline_kind():
    from oracle import make_ref
    return "reference" if os.path.isfile(os.path.join(make_ref.REF_DIR, "PyRayHF", "library.py")) else "port"


def _work(args):
    freq, den, bmag, bpsi, alt, mode, n_points = args
    warnings.simplefilter("ignore")
    import logging
    logging.getLogger("PyRayHF_logger").setLevel(logging.CRITICAL)
    fn, _ = _operator()
    t0 = time.perf_counter()
    vh = fn(freq, den, bmag, bpsi, alt, mode, n_points)
    return time.perf_counter() - t0, vh


def _warm(_):
    _operator()
    return os.getpid()


def one_process(freq, den, bmag, bpsi, alt, mode, n_points, repeats=3):
    """Best-of-``repeats`` wall time of one call in THIS process after one warm-up call -> (seconds, vh)."""
    best, vh = np.inf, None
    for k in range(repeats + 1):
        t, vh = _work((freq, den, bmag, bpsi, alt, mode, n_points))
        if k:
            best = min(best, t)
    return best, vh


class ReferencePool:
    """``cores`` spawned processes; ``one_pass`` hands each ONE profile (all its frequencies) and waits for all."""

    def __init__(self, cores=None):
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_warm, range(self.cores), chunksize=1)          # imports done before anything is timed
        self.kind = baseline_kind()

    def one_pass(self, freq, den, bmag, bpsi, alt, mode, n_points):
        """``den`` / ``bmag`` / ``bpsi`` are ``[k, A]`` with k <= cores profiles.  Returns (wall_s, vh [k, F])."""
        jobs = [(freq, den[q], bmag[q], bpsi[q], alt, mode, n_points) for q in range(den.shape[0])]
        t0 = time.perf_counter()
        res = self.pool.map(_work, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return wall, np.stack([r[1] for r in res])

    def close(self):
        self.pool.close()
        self.pool.join()


def c_port_rate(freq, den, bmag, bpsi, alt, mode, n_points, copies=None):
    """vh/s of the scalar C restatement with every host thread (profiles = independent copies)."""
    from oracle import scalar
    cores = os.cpu_count() or 1
    copies = copies or cores
    d2 = np.tile(den, (copies, 1))
    b2 = np.tile(bmag, (copies, 1))
    p2 = np.tile(bpsi, (copies, 1))
    scalar.vertical_forward_operator_batched(freq[:8], d2[:cores], b2[:cores], p2[:cores], alt, mode, n_points)
    t0 = time.perf_counter()
    scalar.vertical_forward_operator_batched(freq, d2, b2, p2, alt, mode, n_points, n_threads=0)
    wall = time.perf_counter() - t0
    return copies * freq.size / wall, cores

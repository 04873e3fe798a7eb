"""TEST/BENCH INFRASTRUCTURE ONLY -- timing harness for the CPU baseline legs of bench.py.

Runs the numpy restatement of the reference (oracle/vfo_oracle.py, same whole-array numpy
evaluation as PyRayHF/library.py:459-509, bit-identical results) in one process per host core,
because the reference is single-threaded numpy and "all the host threads it can use" means one
independent call per core.  Workers are spawned (never forked from a CUDA process) and import
numpy + the oracle only.
"""
import multiprocessing as mp
import os
import time
import warnings

import numpy as np


def _work(args):
    freq, den, bmag, bpsi, alt, mode, n_points, reps = args
    warnings.simplefilter("ignore")
    from oracle import vfo_oracle
    t0 = time.perf_counter()
    vh = None
    for _ in range(reps):
        vh = vfo_oracle.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points)
    return time.perf_counter() - t0, vh


class NumpyPortPool:
    """A pool of ``cores`` processes, each evaluating the same rows per pass."""

    def __init__(self, cores=None):
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)

    def one_pass(self, freq, den, bmag, bpsi, alt, mode, n_points, reps=1):
        """Every worker evaluates ``freq`` rows ``reps`` times.  Returns (wall_s, rows_done, vh)."""
        job = (freq, den, bmag, bpsi, alt, mode, n_points, reps)
        t0 = time.perf_counter()
        res = self.pool.map(_work, [job] * self.cores, chunksize=1)
        wall = time.perf_counter() - t0
        return wall, self.cores * reps * freq.size, res[0][1]

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

"""CPU, world_size 2 over gloo: the profile-sharding / gather logic of the multi-GPU path.

The compute step is replaced by the oracle (allowed in tests) so that this runs without a GPU;
what is under test is pyrayhf_b200.sharding: every rank passes only its own rows, the result is assembled in the
shared-memory segment (or by a collective gather of padded slices), uneven / empty shards, gather order, and the
agreement on errors (a bad profile in one shard raises everywhere, nobody hangs).
"""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_compute(freq, den, bmag, bpsi, alt, mode, n_points):
    from oracle import scalar, vfo_oracle
    m = vfo_oracle.stretch_multiplier(n_points)
    return scalar.vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode, n_points,
                                                    variant=0, multiplier=m, n_threads=1)[0]


def _inputs(n_prof):
    from pyrayhf_b200 import synth
    lat, lon = synth.grid_subset(n_prof, seed=11)
    alt = synth.default_alt()
    freq = synth.default_freq()[::6]
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    return freq, den, bmag, bpsi, alt


def _worker(rank, world, port, layout, gather, n_prof, result_path):
    sys.path.insert(0, ROOT)
    warnings.simplefilter("ignore")
    import torch.distributed as dist
    from pyrayhf_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        freq, den, bmag, bpsi, alt = _inputs(n_prof)
        idx = sharding.shard_indices(n_prof, world, rank, layout)
        # every rank hands over ITS rows only
        out = sharding.vertical_forward_operator_sharded(freq, den[idx], bmag[idx], bpsi[idx], alt, 'X', 64, layout=layout,
                                                         gather_to=0, gather=gather, compute=_oracle_compute)
        if rank == 0:
            np.save(result_path, out)
        else:
            assert out is None
        # gather on every rank, total count given explicitly; the operator object is reused for two passes
        op = sharding.ShardedForwardOperator(n_prof, freq.size, layout=layout, gather_to=None, gather=gather,
                                             compute=_oracle_compute)
        for _ in range(2):
            out_all = op(freq, den[idx], bmag[idx], bpsi[idx], alt, 'X', 64)
            assert out_all.shape == (n_prof, freq.size)
        first = np.array(out_all, copy=True)
        dist.barrier()
        op.close()
        if rank == 1:
            np.save(result_path + ".rank1.npy", first)
        # a profile the reference rejects (negative density below the peak, library.py:94) sits in ONE shard:
        # every rank must raise the reference's exception, nobody may hang in the gather
        if n_prof >= 2:
            bad = den.copy()
            bad[n_prof - 1, 3] = -1.0
            with pytest.raises(ValueError, match="Density must be non-negative"):
                sharding.vertical_forward_operator_sharded(freq, bad[idx], bmag[idx], bpsi[idx], alt, 'X', 64,
                                                           layout=layout, gather_to=0, gather=gather,
                                                           compute=_oracle_status_compute)
            # a rank that is handed the wrong number of rows: the others must not be left waiting
            with pytest.raises((ValueError, RuntimeError)):
                wrong = idx if rank == 0 else idx[:-1]
                sharding.vertical_forward_operator_sharded(freq, den[wrong], bmag[wrong], bpsi[wrong], alt, 'X', 64,
                                                           n_profiles=n_prof, layout=layout, gather_to=0, gather=gather,
                                                           compute=_oracle_compute)
    finally:
        dist.destroy_process_group()


def _oracle_status_compute(freq, den, bmag, bpsi, alt, mode, n_points):
    """Oracle with the reference's exception for a bad profile (what the CUDA operator reports through status)."""
    from oracle import vfo_oracle
    return np.stack([vfo_oracle.vertical_forward_operator(freq, den[q], bmag[q], bpsi[q], alt, mode, n_points)
                     for q in range(den.shape[0])])


@pytest.mark.parametrize("layout,gather,n_prof", [("interleaved", "shm", 7), ("contiguous", "shm", 7),
                                                   ("interleaved", "collective", 7), ("contiguous", "collective", 5),
                                                   ("interleaved", "shm", 1)])
def test_two_rank_sharding_matches_single_process(tmp_path, layout, gather, n_prof):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 7 * n_prof + len(layout) + 3 * len(gather)) % 2000
    path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(2, port, layout, gather, n_prof, path), nprocs=2, join=True)
    got = np.load(path)
    freq, den, bmag, bpsi, alt = _inputs(n_prof)
    want = _oracle_compute(freq, den, bmag, bpsi, alt, 'X', 64)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(np.load(path + ".rank1.npy"), want, equal_nan=True)


def _oracle_compute_single(freq, den, bmag, bpsi, alt, mode, n_points):
    from oracle import scalar, vfo_oracle
    return scalar.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points, variant=0,
                                            multiplier=vfo_oracle.stretch_multiplier(n_points))


def _freq_worker(rank, world, port, n_freq, result_path):
    sys.path.insert(0, ROOT)
    warnings.simplefilter("ignore")
    import torch.distributed as dist
    from pyrayhf_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        den, bmag, bpsi, alt = synth.single_day_profile()
        freq = np.linspace(0.05, 9.0, n_freq)
        out = sharding.vertical_forward_operator_sharded_by_frequency(freq, den, bmag, bpsi, alt, 'O', 80, gather_to=0,
                                                                      compute=_oracle_compute_single)
        if rank == 0:
            np.save(result_path, out)
        else:
            assert out is None
        out_all = sharding.vertical_forward_operator_sharded_by_frequency(freq, den, bmag, bpsi, alt, 'O', 80,
                                                                          gather_to=None, compute=_oracle_compute_single)
        assert out_all.shape == (n_freq,)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_freq", [37, 1])
def test_two_rank_frequency_sharding_of_one_profile(tmp_path, n_freq):
    """Config 5's shape: one profile, a long frequency sweep split over the ranks (interleaved), gathered on rank 0."""
    import torch.multiprocessing as mp
    from pyrayhf_b200 import synth
    port = 31500 + (os.getpid() + n_freq) % 2000
    path = str(tmp_path / "outf.npy")
    mp.spawn(_freq_worker, args=(2, port, n_freq, path), nprocs=2, join=True)
    got = np.load(path)
    den, bmag, bpsi, alt = synth.single_day_profile()
    freq = np.linspace(0.05, 9.0, n_freq)
    want = _oracle_compute_single(freq, den, bmag, bpsi, alt, 'O', 80)
    assert np.array_equal(got, want, equal_nan=True)


def test_shard_bounds_cover_everything():
    from pyrayhf_b200 import sharding
    for n in (0, 1, 7, 8, 65341):
        for world in (1, 2, 3, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                a, b = sharding.shard_bounds(n, world, r)
                seen[a:b] += 1
                assert abs((b - a) - n / world) < 1.0 + 1e-9
            assert np.all(seen == 1)
            seen[:] = 0
            for r in range(world):
                seen[sharding.interleaved_indices(n, world, r)] += 1
            assert np.all(seen == 1)

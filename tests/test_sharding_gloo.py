"""CPU, world_size 2 over gloo: the profile-sharding / gather logic of the multi-GPU path.

The compute step is replaced by the oracle (allowed in tests) so that this runs without a GPU;
what is under test is pyrayhf_b200.sharding: partitioning, uneven shards, padding, gather order.
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


def _worker(rank, world, port, layout, n_prof, result_path):
    sys.path.insert(0, ROOT)
    warnings.simplefilter("ignore")
    import torch.distributed as dist
    from pyrayhf_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lat, lon = synth.grid_subset(n_prof, seed=11)
        alt = synth.default_alt()
        freq = synth.default_freq()[::6]
        den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
        out = sharding.vertical_forward_operator_sharded(freq, den, bmag, bpsi, alt, 'X', 64, layout=layout,
                                                         gather_to=0, compute=_oracle_compute)
        if rank == 0:
            np.save(result_path, out)
        else:
            assert out is None
        # gather on every rank
        out_all = sharding.vertical_forward_operator_sharded(freq, den, bmag, bpsi, alt, 'X', 64, layout=layout,
                                                             gather_to=None, compute=_oracle_compute)
        assert out_all.shape == (n_prof, freq.size)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout,n_prof", [("interleaved", 7), ("contiguous", 7), ("interleaved", 1)])
def test_two_rank_sharding_matches_single_process(tmp_path, layout, n_prof):
    import torch.multiprocessing as mp
    from pyrayhf_b200 import synth
    port = 29500 + (os.getpid() + n_prof + len(layout)) % 2000
    path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(2, port, layout, n_prof, path), nprocs=2, join=True)
    got = np.load(path)
    lat, lon = synth.grid_subset(n_prof, seed=11)
    alt = synth.default_alt()
    freq = synth.default_freq()[::6]
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    want = _oracle_compute(freq, den, bmag, bpsi, alt, 'X', 64)
    assert np.array_equal(got, want, equal_nan=True)


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

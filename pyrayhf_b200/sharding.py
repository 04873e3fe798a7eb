"""Profile sharding across the GPUs of one node (one process per GPU).

The operator is embarrassingly parallel over profiles (every (profile, frequency) output depends on
that profile's four arrays and one scalar frequency; library.py:459-509 has no cross-profile data
flow), so the multi-GPU form is: split the profile axis, run the single-GPU batched operator on
each shard, gather the [P/G x F] slices.  There is no collective on the data path; the only
communication is the final gather (torch.distributed, NCCL on GPUs / gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_profiles, world_size, rank):
    """Contiguous [start, stop) of `rank`; sizes differ by at most one profile."""
    base, extra = divmod(int(n_profiles), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def interleaved_indices(n_profiles, world_size, rank):
    """Profiles rank, rank + G, rank + 2G, ... (balances the live-row fraction, which follows foF2)."""
    return np.arange(rank, n_profiles, world_size)


def _default_compute(freq, den, bmag, bpsi, alt, mode, n_points):
    from pyrayhf_b200.library import vertical_forward_operator_batched
    return vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode, n_points, errors='nan')


def vertical_forward_operator_sharded(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                      group=None, layout='interleaved', gather_to=0, compute=None):
    """Every rank passes the FULL [P, A] inputs (numpy); each computes its share of the profiles and the
    [P, F] result is assembled on rank `gather_to` (None: on every rank).  Returns the array where it
    is assembled and None elsewhere.

    ``compute`` defaults to the CUDA batched operator of this package; the CPU test-suite passes the
    oracle instead, to exercise the partition / gather logic under gloo without a GPU.
    """
    import torch
    import torch.distributed as dist
    compute = compute or _default_compute
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    den = np.asarray(den)
    n_prof = den.shape[0]
    freq = np.asarray(freq)
    alt = np.asarray(alt)
    n_freq = freq.shape[-1]
    if layout == 'interleaved':
        idx = interleaved_indices(n_prof, world, rank)
    elif layout == 'contiguous':
        idx = np.arange(*shard_bounds(n_prof, world, rank))
    else:
        raise ValueError("layout must be 'interleaved' or 'contiguous'")
    f_loc = freq[idx] if freq.ndim == 2 else freq
    a_loc = alt[idx] if alt.ndim == 2 else alt
    if idx.size:
        local = np.asarray(compute(f_loc, den[idx], np.asarray(bmag)[idx], np.asarray(bpsi)[idx], a_loc,
                                   mode, n_points), dtype=np.float64)
    else:
        local = np.empty((0, n_freq))
    if world == 1:
        out = np.empty((n_prof, n_freq))
        out[idx] = local
        return out
    # pad shards to a common row count so that one all_gather / gather moves them
    rows = -(-n_prof // world)
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    buf = torch.full((rows, n_freq), float('nan'), dtype=torch.float64, device=device)
    if idx.size:
        buf[:idx.size] = torch.from_numpy(local).to(device)
    if gather_to is None or backend == 'nccl':
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == gather_to else None
        dist.gather(buf, parts, dst=gather_to, group=group)
    if gather_to is not None and rank != gather_to:
        return None
    out = np.empty((n_prof, n_freq))
    for rk in range(world):
        ridx = (interleaved_indices(n_prof, world, rk) if layout == 'interleaved'
                else np.arange(*shard_bounds(n_prof, world, rk)))
        out[ridx] = parts[rk][:ridx.size].cpu().numpy()
    return out


def _default_compute_single(freq, den, bmag, bpsi, alt, mode, n_points):
    from pyrayhf_b200.library import vertical_forward_operator
    return vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points)


def vertical_forward_operator_sharded_by_frequency(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                                   group=None, gather_to=0, compute=None):
    """ONE profile, many sounding frequencies (BASELINE config 5: 1 740 of them): every frequency row is independent
    (library.py:459-509 carries nothing from one row to the next), so rank r takes frequencies r, r + G, r + 2G, ...
    -- interleaved, because rows above the critical frequency cost almost nothing and sit at the end of the sweep --
    and the ``[F]`` result is assembled on rank ``gather_to`` (None: on every rank).  Errors of the reference
    (negative density, peak at index 0) surface on every rank, as every rank sees the same profile.
    """
    import torch
    import torch.distributed as dist
    compute = compute or _default_compute_single
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    freq = np.ascontiguousarray(freq, dtype=np.float64).reshape(-1)
    n_freq = freq.size
    idx = interleaved_indices(n_freq, world, rank)
    local = (np.asarray(compute(freq[idx], den, bmag, bpsi, alt, mode, n_points), dtype=np.float64)
             if idx.size else np.empty(0))
    if world == 1:
        out = np.empty(n_freq)
        out[idx] = local
        return out
    rows = -(-n_freq // world)
    backend = dist.get_backend(group)
    device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    buf = torch.full((rows,), float('nan'), dtype=torch.float64, device=device)
    if idx.size:
        buf[:idx.size] = torch.from_numpy(local).to(device)
    if gather_to is None or backend == 'nccl':
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == gather_to else None
        dist.gather(buf, parts, dst=gather_to, group=group)
    if gather_to is not None and rank != gather_to:
        return None
    out = np.empty(n_freq)
    for rk in range(world):
        ridx = interleaved_indices(n_freq, world, rk)
        out[ridx] = parts[rk][:ridx.size].cpu().numpy()
    return out

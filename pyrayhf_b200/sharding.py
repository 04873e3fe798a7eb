"""Profile sharding across the GPUs of one node (one process per GPU).

The operator is embarrassingly parallel over profiles (every (profile, frequency) output depends on that
profile's four arrays and one scalar frequency; library.py:459-509 has no cross-profile data flow), so the
multi-GPU form is: split the profile axis, every rank runs the single-GPU operator on ITS shard
only, and the ``[P/G x F]`` slices are gathered on the host.  There is no collective on the data path:

* each rank holds nothing but its own shard (host arrays, device tensors, or 40 bytes of layer parameters per
  profile from which the profiles are built on the device);
* the gathered ``[P, F]`` result lives in ONE POSIX shared-memory segment that every rank of the node maps and
  page-locks (``cudaHostRegister``); each GPU copies its rows straight into place over its own PCIe link
  (``prhf_vfo_stream_f64`` with a row stride for interleaved shards), overlapped with its kernels;
* the only communication is a barrier (and one tiny all-reduce when an error flag or a profile count has to be
  agreed on).  ``gather='collective'`` keeps a ``torch.distributed.gather`` of padded slices for groups that span
  nodes, where shared memory is not an option.

The CPU test-suite drives the same code under gloo with ``compute=`` the oracle.
"""
import numpy as np

LAYOUTS = ('interleaved', 'contiguous')


def shard_bounds(n_profiles, world_size, rank):
    """Contiguous [start, stop) of `rank`; sizes differ by at most one profile."""
    base, extra = divmod(int(n_profiles), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def interleaved_indices(n_profiles, world_size, rank):
    """Profiles rank, rank + G, rank + 2G, ... (balances the live-row fraction, which follows foF2)."""
    return np.arange(rank, n_profiles, world_size)


def shard_indices(n_profiles, world_size, rank, layout='interleaved'):
    """Global profile indices owned by `rank`."""
    if layout == 'interleaved':
        return interleaved_indices(n_profiles, world_size, rank)
    if layout == 'contiguous':
        return np.arange(*shard_bounds(n_profiles, world_size, rank))
    raise ValueError("layout must be 'interleaved' or 'contiguous'")


def _dist():
    import torch.distributed as dist
    return dist


def _world(group):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _global_rank(group, group_rank):
    """torch.distributed's src / dst arguments are GLOBAL ranks; `group_rank` counts inside `group`."""
    dist = _dist()
    if group is None or group_rank is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


def _comm_device(group):
    import torch
    dist = _dist()
    if dist.get_backend(group) == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def _agree_max(value, group):
    """max of an integer over the ranks of `group` (one 8-byte all-reduce)."""
    import torch
    dist = _dist()
    world, _ = _world(group)
    if world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def _agree_sum(value, group):
    import torch
    dist = _dist()
    world, _ = _world(group)
    if world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def _raise_agreed(code):
    from pyrayhf_b200.library import _raise_profile_status
    _raise_profile_status(code)
    if code:
        raise RuntimeError("the forward operator failed on another rank (status %d)" % code)


_EXC_TYPES = {'ValueError': ValueError, 'IndexError': IndexError, 'TypeError': TypeError}


def _raise_together(failure, exc, group):
    """If any rank of `group` caught an exception before the collective part, every rank raises: the failing
    ranks their own exception, the others the first failing rank's type and message (the reference's exceptions
    are part of the interface: ValueError for a negative density or a bad mode, IndexError for a peak at the
    bottom).  One 8-byte all-reduce in the common case of no failure."""
    if not _agree_max(failure, group):
        return
    world, _ = _world(group)
    if world > 1:
        infos = [None] * world
        _dist().all_gather_object(infos, (type(exc).__name__, str(exc)) if exc is not None else None, group=group)
    if exc is not None:
        raise exc
    name, text = next(i for i in infos if i is not None)
    raise _EXC_TYPES.get(name, RuntimeError)(text)


class HostResult:
    """The gathered result: ``vh`` ``[P, F]`` float64 and ``status`` ``[P]`` int32 in one POSIX shared-memory
    segment mapped by every rank of ``group`` (ranks of one node) and page-locked in every rank that has a GPU, so
    that each GPU's copy engine writes its rows in place.  World size 1: plain pinned memory."""

    def __init__(self, n_profiles, n_freq, group=None, owner=0, register=None):
        import torch
        self.n_profiles, self.n_freq = int(n_profiles), int(n_freq)
        self.group = group
        self.world, self.rank = _world(group)
        self._shm = None
        self._registered = None
        vh_bytes = self.n_profiles * self.n_freq * 8
        total = max(vh_bytes + self.n_profiles * 4, 8)
        if register is None:
            register = torch.cuda.is_available()
        if self.world == 1:
            from pyrayhf_b200.library import pinned_empty
            raw = pinned_empty(total, dtype=np.uint8)
        else:
            # a file in /dev/shm mapped by every rank; the owner unlinks it as soon as everybody has it open, so
            # nothing is left behind even if a rank dies (the mappings keep the memory alive)
            import mmap
            import os
            import secrets
            dist = _dist()
            name = [None]
            fd = -1
            if self.rank == owner:
                name[0] = "/dev/shm/prhf_%d_%s" % (os.getpid(), secrets.token_hex(6))
                fd = os.open(name[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, total)
            dist.broadcast_object_list(name, src=_global_rank(group, owner), group=group)
            if self.rank != owner:
                fd = os.open(name[0], os.O_RDWR)
            self._shm = mmap.mmap(fd, total)
            os.close(fd)
            dist.barrier(group=group)
            if self.rank == owner:
                os.unlink(name[0])
            raw = np.frombuffer(self._shm, dtype=np.uint8)
            if register:
                from pyrayhf_b200 import _cabi
                _cabi.host_register(raw.ctypes.data, total)
                self._registered = (raw.ctypes.data, total)
        self._raw = raw
        self.vh = raw[:vh_bytes].view(np.float64).reshape(self.n_profiles, self.n_freq)
        self.status = raw[vh_bytes:vh_bytes + self.n_profiles * 4].view(np.int32)

    def close(self):
        if self._registered is not None:
            from pyrayhf_b200 import _cabi
            _cabi.host_unregister(self._registered[0])
            self._registered = None
        self.vh = self.status = self._raw = None
        if self._shm is not None:
            shm, self._shm = self._shm, None
            try:
                shm.close()
            except BufferError:          # a caller still holds a view: the mapping goes with the last reference
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedForwardOperator:
    """``vertical_forward_operator`` for ``n_profiles`` profiles split over the ranks of ``group``.

    Rank ``r`` owns the profiles ``shard_indices(n_profiles, G, r, layout)`` and passes ONLY those (``[P_r, A]``
    arrays, host or device, or ``[P_r, 5]`` layer parameters to ``from_parameters``).  Every call is collective
    over the group.  The ``[n_profiles, n_freq]`` result is returned on rank ``gather_to`` (on every rank when
    ``gather_to`` is None) as a view of the shared result buffer -- valid until the next call or ``close()`` --
    and ``None`` elsewhere.  Per-profile errors of the reference (negative density, peak at the bottom) are agreed
    between the ranks before anyone raises, so a bad profile on one shard cannot leave the others in a barrier.

    ``compute`` (tests): ``compute(freq, den, bmag, bpsi, alt, mode, n_points) -> [P_r, F]`` replaces the CUDA
    operator, e.g. with the oracle under gloo.
    """

    def __init__(self, n_profiles, n_freq, *, group=None, layout='interleaved', gather_to=0, gather='shm',
                 compute=None, chunk_profiles=0):
        if layout not in LAYOUTS:
            raise ValueError("layout must be 'interleaved' or 'contiguous'")
        if gather not in ('shm', 'collective'):
            raise ValueError("gather must be 'shm' or 'collective'")
        self.n_profiles, self.n_freq = int(n_profiles), int(n_freq)
        self.group, self.layout, self.gather_to, self.gather = group, layout, gather_to, gather
        self.compute = compute
        self.chunk_profiles = chunk_profiles
        self.world, self.rank = _world(group)
        if gather_to is not None and not (0 <= gather_to < self.world):
            raise ValueError("gather_to must be a rank of the group or None")
        self.indices = shard_indices(self.n_profiles, self.world, self.rank, layout)
        self.n_local = int(self.indices.size)
        self.result = None
        if gather == 'shm' or self.world == 1:
            self.result = HostResult(self.n_profiles, self.n_freq, group,
                                     owner=0 if gather_to is None else gather_to,
                                     register=None if compute is None else False)

    # -- where this rank's rows go inside the gathered array --
    def _local_out(self):
        vh = self.result.vh
        if self.layout == 'contiguous' or self.world == 1:
            start = int(self.indices[0]) if self.n_local else 0
            return vh[start:start + self.n_local], self.n_freq
        flat = vh.reshape(-1)
        return flat[self.rank * self.n_freq:], self.world * self.n_freq

    def _finish(self, local_status, errors, local_vh=None):
        """Agree on errors, complete the gather, hand the result to the ranks that asked for it."""
        dist = _dist()
        worst_local = int(np.max(local_status)) if self.n_local else 0
        if self.result is not None:
            if self.n_local:
                self.result.status[self.indices] = local_status
            if self.world > 1:
                dist.barrier(group=self.group)               # every rank's rows and status have landed
            worst = int(self.result.status.max()) if self.n_profiles else 0
            if errors == 'raise' and worst:
                _raise_agreed(worst)
            wanted = self.gather_to is None or self.rank == self.gather_to
            return self.result.vh if wanted else None
        # collective gather of padded slices (groups that span nodes)
        import torch
        worst = _agree_max(worst_local, self.group)
        if errors == 'raise' and worst:
            _raise_agreed(worst)
        rows = -(-self.n_profiles // self.world)
        dev = _comm_device(self.group)
        buf = torch.full((rows, self.n_freq), float('nan'), dtype=torch.float64, device=dev)
        if self.n_local:
            src = local_vh if hasattr(local_vh, 'data_ptr') else torch.from_numpy(np.ascontiguousarray(local_vh))
            buf[:self.n_local] = src.to(dev)
        if self.gather_to is None:
            parts = [torch.empty_like(buf) for _ in range(self.world)]
            dist.all_gather(parts, buf, group=self.group)
        else:
            parts = [torch.empty_like(buf) for _ in range(self.world)] if self.rank == self.gather_to else None
            dist.gather(buf, parts, dst=_global_rank(self.group, self.gather_to), group=self.group)
            if self.rank != self.gather_to:
                return None
        out = np.empty((self.n_profiles, self.n_freq))
        for rk in range(self.world):
            ridx = shard_indices(self.n_profiles, self.world, rk, self.layout)
            out[ridx] = parts[rk][:ridx.size].cpu().numpy()
        return out

    def __call__(self, freq, den, bmag, bpsi, alt, mode='O', n_points=200, *, errors='raise', stream=None):
        """One sharded pass.  ``den`` / ``bmag`` / ``bpsi`` are this rank's ``[P_r, A]`` rows; ``freq`` ``[F]`` (or this
        rank's ``[P_r, F]``); ``alt`` ``[A]`` (or ``[P_r, A]``)."""
        local_status = np.zeros(self.n_local, dtype=np.int32)
        local_vh = None
        failure, local_exc = 0, None
        try:
            if int(den.shape[0]) != self.n_local:
                raise ValueError("rank %d owns %d of the %d profiles (layout %r) but was given %d"
                                 % (self.rank, self.n_local, self.n_profiles, self.layout, int(den.shape[0])))
            if int(freq.shape[-1]) != self.n_freq:
                raise ValueError("freq has %d frequencies, the operator was built for %d"
                                 % (int(freq.shape[-1]), self.n_freq))
            if self.compute is not None:
                local_vh = np.empty((0, self.n_freq))
                if self.n_local:
                    local_vh = np.asarray(self.compute(freq, den, bmag, bpsi, alt, mode, n_points), dtype=np.float64)
                if self.result is not None and self.n_local:
                    self.result.vh[self.indices] = local_vh
            elif self.result is not None:
                from pyrayhf_b200.library import vertical_forward_operator_streamed
                if self.n_local:
                    out, stride = self._local_out()
                    vertical_forward_operator_streamed(freq, den, bmag, bpsi, alt, mode, n_points, errors='nan',
                                                       out=out, out_profile_stride=stride, status_out=local_status,
                                                       chunk_profiles=self.chunk_profiles, stream=stream)
            else:
                from pyrayhf_b200.library import vertical_forward_operator_streamed
                import torch
                dev = torch.device('cuda', torch.cuda.current_device())
                local_vh = torch.empty((self.n_local, self.n_freq), dtype=torch.float64, device=dev)
                if self.n_local:
                    vertical_forward_operator_streamed(freq, den, bmag, bpsi, alt, mode, n_points, errors='nan',
                                                       out=local_vh, status_out=local_status,
                                                       chunk_profiles=self.chunk_profiles, stream=stream)
        except (ValueError, IndexError, TypeError) as exc:     # every rank must leave the collective part together
            failure, local_exc = 1, exc
        _raise_together(failure, local_exc, self.group)
        return self._finish(local_status, errors, local_vh)

    def from_parameters(self, freq, params, alt, mode='O', n_points=200, *, errors='raise'):
        """Profiles built ON THE DEVICE from this rank's ``[P_r, 5]`` layer parameters {foF2 MHz, hmF2 km, scale height
        km, foE MHz, latitude deg} (``prhf_synth_profiles_f64``: two Chapman layers + dipole field), then the sharded
        pass: 40 bytes per profile cross PCIe instead of 15 KB (BASELINE configs[3], SURVEY.md 8d "Config 4")."""
        import torch
        from pyrayhf_b200 import synth
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 5)
        dev = torch.device('cuda', torch.cuda.current_device())
        den, bmag, bpsi = synth.profiles_from_parameters_device(*params.T, alt=alt, device=dev)
        t_alt = torch.from_numpy(np.ascontiguousarray(alt, dtype=np.float64)).to(dev)
        t_freq = torch.from_numpy(np.ascontiguousarray(freq, dtype=np.float64)).to(dev)
        return self(t_freq, den, bmag, bpsi, t_alt, mode, n_points, errors=errors,
                    stream=torch.cuda.current_stream(dev).cuda_stream)

    def close(self):
        if self.result is not None:
            self.result.close()
            self.result = None


def vertical_forward_operator_sharded(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *, n_profiles=None,
                                      group=None, layout='interleaved', gather_to=0, gather='shm', compute=None,
                                      errors='raise'):
    """One-shot form of ``ShardedForwardOperator``: every rank passes ITS rows only (``[P_r, A]``, the profiles
    ``shard_indices(P, G, rank, layout)`` of the global batch); ``n_profiles`` defaults to the sum of the local
    counts.  Returns a new ``[P, F]`` array on rank ``gather_to`` (every rank when None), ``None`` elsewhere."""
    n_local = int(den.shape[0])
    total = _agree_sum(n_local, group) if n_profiles is None else int(n_profiles)
    op = ShardedForwardOperator(total, int(freq.shape[-1]), group=group, layout=layout, gather_to=gather_to,
                                gather=gather, compute=compute)
    try:
        out = op(freq, den, bmag, bpsi, alt, mode, n_points, errors=errors)
        return None if out is None else np.array(out, copy=True)
    finally:
        world, _ = _world(group)
        if world > 1 and op.result is not None:
            _dist().barrier(group=group)                     # nobody unlinks the segment while a peer still copies
        op.close()


def _default_compute_single(freq, den, bmag, bpsi, alt, mode, n_points):
    from pyrayhf_b200.library import vertical_forward_operator
    return vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points)


def vertical_forward_operator_sharded_by_frequency(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                                   group=None, gather_to=0, compute=None):
    """ONE profile, many sounding frequencies (BASELINE config 5: 1 740 of them): every frequency row is independent
    (library.py:459-509 carries nothing from one row to the next), so rank r takes frequencies r, r + G, r + 2G, ...
    -- interleaved, because rows above the critical frequency cost almost nothing and sit at the end of the sweep --
    and the ``[F]`` result is assembled on rank ``gather_to`` (None: on every rank).  Every rank sees the same
    profile, so the reference's errors (negative density, peak at index 0, bad mode) are raised on every rank: a rank
    without frequencies of its own (G > F) still validates the profile on one frequency, and the outcome is agreed
    with one all-reduce before anyone enters the gather.
    """
    import torch
    dist = _dist()
    compute = compute or _default_compute_single
    world, rank = _world(group)
    freq = np.ascontiguousarray(freq, dtype=np.float64).reshape(-1)
    n_freq = freq.size
    idx = interleaved_indices(n_freq, world, rank)
    local = np.empty(0)
    failure, local_exc = 0, None
    try:
        if idx.size:
            local = np.asarray(compute(freq[idx], den, bmag, bpsi, alt, mode, n_points), dtype=np.float64)
        elif n_freq:
            compute(freq[:1], den, bmag, bpsi, alt, mode, n_points)      # validation only
    except (ValueError, IndexError, TypeError) as exc:
        failure, local_exc = 1, exc
    _raise_together(failure, local_exc, group)
    if world == 1:
        out = np.empty(n_freq)
        out[idx] = local
        return out
    rows = -(-n_freq // world)
    device = _comm_device(group)
    buf = torch.full((rows,), float('nan'), dtype=torch.float64, device=device)
    if idx.size:
        buf[:idx.size] = torch.from_numpy(local).to(device)
    if gather_to is None:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == gather_to else None
        dist.gather(buf, parts, dst=_global_rank(group, gather_to), group=group)
        if rank != gather_to:
            return None
    out = np.empty(n_freq)
    for rk in range(world):
        ridx = interleaved_indices(n_freq, world, rk)
        out[ridx] = parts[rk][:ridx.size].cpu().numpy()
    return out

"""Host-side mirror of the reference interface for the vertical forward operator.

``vertical_forward_operator`` keeps the signature, units, return type and error behaviour
of ``PyRayHF.library.vertical_forward_operator`` (PyRayHF/library.py:459-509); the
arithmetic runs in the fused CUDA kernels behind the C ABI (``include/pyrayhf_b200.h``).
``vertical_forward_operator_batched`` is the ``[n_profiles x n_alt]`` form.

There is no CPU fallback: a missing extension or GPU raises.
"""
import ctypes

import numpy as np

from pyrayhf_b200 import _cabi

_vp = ctypes.c_void_p


def _logger():
    import pyrayhf_b200
    return pyrayhf_b200.logger


def _mode_code(mode):  # noqa: E302
    # library.py:391-396: exact, case-sensitive comparison with 'O' / 'X'
    if isinstance(mode, str) and mode == 'O':
        return 0
    if isinstance(mode, str) and mode == 'X':
        return 1
    raise ValueError("mode must be 'O' or 'X'")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return _vp(a.ctypes.data)


def _raise_profile_status(st):
    if st == 1:
        raise ValueError("Density must be non-negative")                     # library.py:94
    if st == 2:
        raise IndexError("index -1 is out of bounds for axis 1 with size 0")  # library.py:399


_F64 = np.dtype(np.float64)
_status_buf = {}


def _vec(a):
    """float64, C-contiguous, 1-D view/copy of ``a`` (fast path: already in that form)."""
    if type(a) is np.ndarray and a.dtype == _F64 and a.ndim == 1 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def vertical_forward_operator(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                              literal=False, device=-1):
    """Calculate virtual height from ionosonde freq and ion profile (drop-in).

    Parameters as the reference (library.py:463-478): ``freq`` [MHz] ndarray, ``den``
    [m^-3], ``bmag`` [T], ``bpsi`` [deg], ``alt`` [km] ndarrays of one length, ``mode``
    'O' or 'X', ``n_points`` grid points.  Returns a new float64 ndarray, NaN where the
    ray does not reflect.  Raises what the reference raises: ValueError for a bad mode
    (library.py:396) or negative density below the peak (library.py:94), IndexError when
    the density peak is the first sample (library.py:399).

    ``literal=True`` evaluates the Appleton-Hartree block in the reference's operation
    order (debug aid; O-mode then inherits the reference's cancellation noise).
    """
    if mode == 'O':                         # library.py:391-396: exact, case-sensitive
        code = 0
    elif mode == 'X':
        code = 1
    else:
        raise ValueError("mode must be 'O' or 'X'")
    n_points = int(n_points)
    if type(freq) is not np.ndarray:
        # what `freq * 1e6` (library.py:491) and `f.size` (library.py:367) do to non-arrays in the reference
        if isinstance(freq, (list, tuple)):
            raise TypeError("can't multiply sequence by non-int of type 'float'")
        if isinstance(freq, (int, float)) and not isinstance(freq, np.generic):
            raise AttributeError("'float' object has no attribute 'size'")
    ctx = _cabi.context(device)
    st = _status_buf.get(ctx)
    if st is None:
        st = _status_buf[ctx] = np.zeros(1, dtype=np.int32)
    flags = _cabi.FLAG_LITERAL if literal else 0
    fast = ctx.fast
    if fast is not None and n_points >= 1 and type(freq) is np.ndarray and freq.ndim == 1:
        # float64 contiguous vectors go straight to the C ABI through the buffer-protocol shim
        vh = np.empty(freq.shape[0], dtype=np.float64)
        rc = fast.vfo_host(ctx.fn_addr, ctx.ctx_addr, freq, den, bmag, bpsi, alt, code, n_points, flags, vh, st)
        if rc == 0:
            if st[0]:
                _raise_profile_status(int(st[0]))
            return vh
        if rc > 0:
            ctx.check(rc)
    if np.ndim(freq) > 1 and not (np.ndim(freq) == 2 and np.shape(freq)[0] == 1):
        # the reference broadcasts freq against [n_alt, n_freq]: 0-d, 1-d and (1, F) work, anything else fails
        raise ValueError("operands could not be broadcast together: freq must be 0-d, 1-d or of shape (1, F)")
    f, d, b, p, a = _vec(freq), _vec(den), _vec(bmag), _vec(bpsi), _vec(alt)
    n_alt = d.size
    if not (b.size == p.size == a.size == n_alt):
        # library.py:487-488 only logs a shape mismatch and then fails inside numpy; there is no
        # meaningful result to reproduce for ragged inputs
        _logger().error("Error: freq, den, bmag, bpsi, alt should have same size")
        raise ValueError("den, bmag, bpsi, alt must have the same length")
    if n_alt == 0:
        raise ValueError("attempt to get argmax of an empty sequence")       # np.argmax, library.py:371
    if f.size == 0:
        # np.apply_along_axis at library.py:403 refuses a zero-length frequency axis
        raise ValueError("Cannot apply_along_axis when any iteration dimensions are 0")
    if n_points < 1:
        # np.nanmax of an empty [F x 0] array at library.py:201
        raise ValueError("zero-size array to reduction operation fmax which has no identity")
    vh = np.empty(f.size, dtype=np.float64)
    rc = ctx.vfo_host(f.ctypes.data, f.size, 0, d.ctypes.data, b.ctypes.data, p.ctypes.data, a.ctypes.data, 0,
                      1, n_alt, code, n_points, flags, vh.ctypes.data, st.ctypes.data)
    if rc:
        ctx.check(rc)
    if st[0]:
        _raise_profile_status(int(st[0]))
    return vh


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _n_points_checked(n_points):
    n_points = int(n_points)
    if n_points < 1:
        # np.nanmax of an empty [F x 0] array at library.py:201 -- same error as the single-profile entry
        raise ValueError("zero-size array to reduction operation fmax which has no identity")
    return n_points


def _check_batched_shapes(freq, den, bmag, bpsi, alt):
    """Shapes the batched entries accept: den [P, A]; bmag / bpsi [P, A] or [A]; freq [F], (1, F) or [P, F];
    alt [A] or [P, A].  Returns (n_prof, n_alt, n_freq).  Raw pointers go to the C ABI after this, so every
    mismatch must be caught here (a short buffer would be read out of bounds)."""
    if den.ndim != 2:
        raise ValueError("den must be [n_profiles, n_alt], got shape %s" % (tuple(den.shape),))
    n_prof, n_alt = int(den.shape[0]), int(den.shape[1])
    for name, v in (("bmag", bmag), ("bpsi", bpsi)):
        if tuple(v.shape) not in ((n_prof, n_alt), (n_alt,)):
            raise ValueError("%s must be [n_profiles, n_alt] = %s or [n_alt], got %s"
                             % (name, (n_prof, n_alt), tuple(v.shape)))
    if tuple(alt.shape) not in ((n_prof, n_alt), (n_alt,)):
        raise ValueError("alt must be [n_alt] = (%d,) or [n_profiles, n_alt], got %s" % (n_alt, tuple(alt.shape)))
    if freq.ndim == 1:
        n_freq = int(freq.shape[0])
    elif freq.ndim == 2 and int(freq.shape[0]) in (1, n_prof):
        n_freq = int(freq.shape[1])
    else:
        raise ValueError("freq must be [n_freq], (1, n_freq) or [n_profiles, n_freq], got %s" % (tuple(freq.shape),))
    return n_prof, n_alt, n_freq


def _check_out(out, n_prof, n_freq, want_torch, device=None):
    import_ok = _is_torch_tensor(out) if want_torch else isinstance(out, np.ndarray)
    if not import_ok:
        raise ValueError("out must be a %s" % ("torch tensor on the inputs' device" if want_torch else "numpy array"))
    if tuple(out.shape) != (n_prof, n_freq):
        raise ValueError("out must have shape %s, got %s" % ((n_prof, n_freq), tuple(out.shape)))
    if want_torch:
        import torch
        if out.dtype != torch.float64 or not out.is_contiguous() or out.device != device:
            raise ValueError("out must be a contiguous float64 tensor on %s" % (device,))
    elif out.dtype != _F64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array")


def vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                      literal=False, errors='raise', return_status=False,
                                      out=None, device=None):
    """Batched operator: ``den`` is ``[P, A]``; ``bmag`` / ``bpsi`` are ``[P, A]`` or ``[A]`` (shared by all
    profiles); ``freq`` is ``[F]`` or ``[P, F]``; ``alt`` is ``[A]`` or ``[P, A]``.  Row ``p`` of the ``[P, F]``
    result equals the reference called on profile ``p`` alone (peak truncation, unmagnetised switch and
    error status are per profile).  Anything else raises ``ValueError`` before the C ABI sees a pointer.

    Inputs may be numpy arrays (copied through pinned memory, numpy result) or float64
    CUDA torch tensors (used in place on the current stream, torch result, asynchronous).
    ``errors='raise'`` re-raises the reference's per-profile exceptions (needs a host
    sync for torch inputs); ``errors='nan'`` leaves failed profiles as NaN rows.

    A ctx serves one launch sequence at a time: calls issued from one host thread on different torch streams
    are ordered on the device (the later call's stream waits for the earlier one), they do not overlap.
    """
    code = _mode_code(mode)
    n_points = _n_points_checked(n_points)
    flags = _cabi.FLAG_LITERAL if literal else 0
    if _is_torch_tensor(den):
        import torch
        ts = [freq, den, bmag, bpsi, alt]
        for t in ts:
            if not (_is_torch_tensor(t) and t.is_cuda and t.dtype == torch.float64 and t.device == den.device):
                raise TypeError("torch inputs must all be float64 CUDA tensors on one device")
        n_prof, n_alt, n_freq = _check_batched_shapes(*ts)
        if bmag.dim() == 1:
            bmag = bmag.expand(n_prof, n_alt)
        if bpsi.dim() == 1:
            bpsi = bpsi.expand(n_prof, n_alt)
        if freq.dim() == 2 and freq.shape[0] == 1:
            freq = freq.reshape(-1)
        freq, den, bmag, bpsi, alt = (t.contiguous() for t in (freq, den, bmag, bpsi, alt))
        dev = den.device
        if out is not None:
            _check_out(out, n_prof, n_freq, True, dev)
        vh = out if out is not None else torch.empty((n_prof, n_freq), dtype=torch.float64, device=dev)
        st = torch.zeros(n_prof, dtype=torch.int32, device=dev)
        if n_prof == 0 or n_freq == 0:
            return (vh, st) if return_status else vh
        ctx = _cabi.context(dev.index if dev.index is not None else torch.cuda.current_device())
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = ctx.lib.prhf_vfo_f64(ctx.handle, _vp(freq.data_ptr()), n_freq, n_freq if freq.dim() == 2 else 0,
                                  _vp(den.data_ptr()), _vp(bmag.data_ptr()), _vp(bpsi.data_ptr()),
                                  _vp(alt.data_ptr()), n_alt if alt.dim() == 2 else 0, n_prof, n_alt, code,
                                  n_points, flags, _vp(vh.data_ptr()), _vp(st.data_ptr()), _vp(stream))
        ctx.check(rc)
        if errors == 'raise':
            bad = int(st.max().item())
            if bad:
                _raise_profile_status(bad)
        return (vh, st) if return_status else vh

    freq, den, bmag, bpsi, alt = (np.asarray(v, dtype=np.float64) for v in (freq, den, bmag, bpsi, alt))
    n_prof, n_alt, n_freq = _check_batched_shapes(freq, den, bmag, bpsi, alt)
    if bmag.ndim == 1:
        bmag = np.broadcast_to(bmag, den.shape)
    if bpsi.ndim == 1:
        bpsi = np.broadcast_to(bpsi, den.shape)
    if freq.ndim == 2 and freq.shape[0] == 1:
        freq = freq.reshape(-1)
    freq, den, bmag, bpsi, alt = (np.ascontiguousarray(v) for v in (freq, den, bmag, bpsi, alt))
    if out is not None:
        _check_out(out, n_prof, n_freq, False)
    vh = out if out is not None else np.empty((n_prof, n_freq), dtype=np.float64)
    st = np.zeros(n_prof, dtype=np.int32)
    if n_prof and n_freq:
        ctx = _cabi.context(-1 if device is None else device)
        rc = ctx.lib.prhf_vfo_host_f64(ctx.handle, _ptr(freq), n_freq, n_freq if freq.ndim == 2 else 0,
                                       _ptr(den), _ptr(bmag), _ptr(bpsi), _ptr(alt),
                                       n_alt if alt.ndim == 2 else 0, n_prof, n_alt, code, n_points,
                                       flags, _ptr(vh), _ptr(st))
        ctx.check(rc)
    if errors == 'raise' and st.any():
        _raise_profile_status(int(st.max()))
    return (vh, st) if return_status else vh


def pinned_empty(shape, dtype=np.float64):
    """A page-locked numpy array (backed by a pinned torch tensor): the input / output memory the streaming
    entry copies from and to at full PCIe rate without a staging pass on the host.  (Without a CUDA device --
    host-side tests of the sharding logic -- a plain array is returned.)"""
    import torch
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(v) for v in shape)
    if not torch.cuda.is_available():
        return np.empty(shape, dtype=dtype)
    tdtype = torch.from_numpy(np.empty(0, dtype=dtype)).dtype
    return torch.empty(shape, dtype=tdtype, pin_memory=True).numpy()   # .base keeps the pinned allocation alive


def _addr_of(x):
    return _vp(x.data_ptr()) if _is_torch_tensor(x) else _vp(x.ctypes.data)


def vertical_forward_operator_streamed(freq, den, bmag, bpsi, alt, mode='O', n_points=200, *,
                                       literal=False, errors='raise', return_status=False, out=None,
                                       out_profile_stride=None, status_out=None, chunk_profiles=0, device=None,
                                       stream=None, synchronize=True):
    """The batched operator for LARGE batches whose arrays live on the host (ideally page-locked, see
    ``pinned_empty``), on the GPU, or a mix: profiles are processed in chunks and the host-to-device copy of the
    next chunk, the kernels of the current one and the device-to-host copy of the previous one overlap
    (``prhf_vfo_stream_f64``).  Shapes as ``vertical_forward_operator_batched`` (no ``[A]`` broadcast of
    ``bmag`` / ``bpsi`` here: pass ``[P, A]``).  float64, C-contiguous numpy arrays or torch tensors (CPU or CUDA).

    ``out`` may be given (numpy / torch, host or device); ``out_profile_stride`` (in doubles) lets the rows land
    strided in a larger buffer -- the sharded operator passes a slice of the gathered result.  Returns ``out``
    (a new pinned numpy array when not given).  With ``synchronize=False`` the call returns once everything is
    enqueued on ``stream`` (a raw CUDA stream handle, default: the ctx's own) and ``errors`` must be ``'nan'``.
    """
    code = _mode_code(mode)
    n_points = _n_points_checked(n_points)
    flags = _cabi.FLAG_LITERAL if literal else 0
    arrs = [freq, den, bmag, bpsi, alt]
    for k, v in enumerate(arrs):
        if _is_torch_tensor(v):
            import torch
            if v.dtype != torch.float64 or not v.is_contiguous():
                raise TypeError("tensors must be contiguous float64")
        else:
            arrs[k] = np.ascontiguousarray(v, dtype=np.float64)
    freq, den, bmag, bpsi, alt = arrs
    n_prof, n_alt, n_freq = _check_batched_shapes(freq, den, bmag, bpsi, alt)
    if tuple(bmag.shape) != (n_prof, n_alt) or tuple(bpsi.shape) != (n_prof, n_alt):
        raise ValueError("the streaming entry takes bmag / bpsi as [n_profiles, n_alt]")
    if freq.ndim == 2 and freq.shape[0] == 1 and n_prof != 1:
        freq = freq.reshape(-1)
    if device is None:
        for v in arrs + [out]:
            if _is_torch_tensor(v) and v.is_cuda:
                device = v.device.index
                break
    stride = n_freq if out_profile_stride is None else int(out_profile_stride)
    if out is None:
        if stride != n_freq:
            raise ValueError("out_profile_stride needs an explicit out buffer")
        out = pinned_empty((n_prof, n_freq))
    elif stride == n_freq:
        if tuple(out.shape) != (n_prof, n_freq):
            raise ValueError("out must have shape %s, got %s" % ((n_prof, n_freq), tuple(out.shape)))
    if _is_torch_tensor(out):
        import torch
        if out.dtype != torch.float64 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float64 tensor")
        out_elems = out.numel()
    else:
        if out.dtype != _F64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array")
        out_elems = out.size
    if n_prof and out_elems < (n_prof - 1) * stride + n_freq:
        raise ValueError("out is too small for %d profiles at a stride of %d doubles" % (n_prof, stride))
    st = status_out if status_out is not None else np.zeros(n_prof, dtype=np.int32)
    if not synchronize and errors == 'raise':
        raise ValueError("errors='raise' needs synchronize=True")
    if n_prof and n_freq:
        ctx = _cabi.context(-1 if device is None else device)
        rc = ctx.lib.prhf_vfo_stream_f64(ctx.handle, _addr_of(freq), n_freq, n_freq if freq.ndim == 2 else 0,
                                         _addr_of(den), _addr_of(bmag), _addr_of(bpsi), _addr_of(alt),
                                         n_alt if alt.ndim == 2 else 0, n_prof, n_alt, code, n_points, flags,
                                         int(chunk_profiles), _addr_of(out), stride, _addr_of(st),
                                         _vp(stream) if stream else None, 1 if synchronize else 0)
        ctx.check(rc)
    if errors == 'raise':
        bad = int(st.max()) if n_prof else 0
        if bad:
            _raise_profile_status(bad)
    return (out, st) if return_status else out


def find_mu_mup(X, Y, bpsi, mode, y_tol=1e-12, *, literal=False):
    """Phase / group refractive index on the GPU (library.py:161-256), numpy in/out."""
    X = np.asarray(X, dtype=float)
    Y = np.asarray(Y, dtype=float)
    bpsi = np.asarray(bpsi, dtype=float)
    X, Y, bpsi = np.broadcast_arrays(X, Y, bpsi)
    shape = X.shape
    with np.errstate(all='ignore'):
        iso = bool(np.nanmax(np.abs(Y)) < y_tol) if Y.size else False      # library.py:201
    if not iso and mode not in ('O', 'X'):
        raise ValueError("Mode must be O or X")                              # library.py:226
    code = 0 if mode == 'O' else 1
    import torch
    dev = torch.device('cuda', torch.cuda.current_device())
    tx, ty, tp = (torch.from_numpy(_f64(v).reshape(-1)).to(dev) for v in (X, Y, bpsi))
    mu = torch.empty_like(tx)
    mup = torch.empty_like(tx)
    ctx = _cabi.context(dev.index)
    rc = ctx.lib.prhf_mu_mup_f64(ctx.handle, _vp(tx.data_ptr()), _vp(ty.data_ptr()), _vp(tp.data_ptr()),
                                 tx.numel(), code, int(iso), _cabi.FLAG_LITERAL if literal else 0,
                                 _vp(mu.data_ptr()), _vp(mup.data_ptr()),
                                 _vp(torch.cuda.current_stream(dev).cuda_stream))
    ctx.check(rc)
    return mu.cpu().numpy().reshape(shape), mup.cpu().numpy().reshape(shape)


def residual_VH_batched(vh_obs, vh_model, *, return_residual=True):
    """Residuals of a batch of modelled virtual-height curves against one observed curve.

    The arithmetic tail of ``residual_VH`` (library.py:660-668) for ``vh_model`` of shape ``[P, F]``:
    NaN model heights become ``max(nanmean(|vh_model[p]|), 100)`` (library.py:664-665) and
    ``residual[p] = vh_obs - vh_model[p]`` (library.py:668).  Returns ``(residual [P, F], chi2 [P])`` with
    ``chi2 = sum(residual**2, axis=1)``, the quantity the brute-force search minimises
    (library.py:794-798).  float64 CUDA tensors in -> tensors out (asynchronous); numpy in -> numpy out.
    """
    import torch
    is_t = _is_torch_tensor(vh_model)
    if is_t:
        dev = vh_model.device
        vm = vh_model.contiguous()
        vo = vh_obs.to(dev).contiguous() if _is_torch_tensor(vh_obs) else torch.from_numpy(_f64(vh_obs)).to(dev)
    else:
        dev = torch.device('cuda', torch.cuda.current_device())
        vm = torch.from_numpy(_f64(vh_model)).to(dev)
        vo = torch.from_numpy(_f64(vh_obs).reshape(-1)).to(dev)
    if vm.dim() != 2 or vo.numel() != vm.shape[1]:
        raise ValueError("vh_model must be [P, F] and vh_obs [F]")
    n_prof, n_freq = vm.shape
    res = torch.empty_like(vm) if return_residual else None
    chi2 = torch.empty(n_prof, dtype=torch.float64, device=dev)
    ctx = _cabi.context(dev.index)
    rc = ctx.lib.prhf_residual_f64(ctx.handle, _vp(vm.data_ptr()), _vp(vo.data_ptr()), n_prof, n_freq,
                                   _vp(res.data_ptr()) if res is not None else None, _vp(chi2.data_ptr()),
                                   _vp(torch.cuda.current_stream(dev).cuda_stream))
    ctx.check(rc)
    if is_t:
        return res, chi2
    return (res.cpu().numpy() if res is not None else None), chi2.cpu().numpy()


_saved = {}


_STAGE_NAMES = ('den2freq', 'find_X', 'find_Y', 'smooth_nonuniform_grid', 'regrid_to_nonuniform_grid',
                'find_mu_mup', 'find_vh')


def install(stages=False):
    """Rebind ``PyRayHF.library.vertical_forward_operator`` to this implementation.

    ``model_VH`` resolves the name through its module globals at call time
    (library.py:589), so the inversion code picks the GPU path up unchanged.
    ``stages=True`` also rebinds the standalone stage functions (pyrayhf_b200/stages.py); leave it off
    when the reference's ray tracers are in use -- they call ``find_X`` / ``find_mu_mup`` on a handful of
    values per step, where a GPU round trip per call is slower than numpy.
    """
    import PyRayHF.library as ref
    names = ('vertical_forward_operator',) + (_STAGE_NAMES if stages else ())
    import pyrayhf_b200
    for name in names:
        if name not in _saved:
            _saved[name] = getattr(ref, name)
        setattr(ref, name, getattr(pyrayhf_b200, name))
    return ref


def uninstall():
    if _saved:
        import PyRayHF.library as ref
        for name in list(_saved):
            setattr(ref, name, _saved.pop(name))

"""The stages of the vertical-forward-operator path as standalone GPU operators.

``vertical_forward_operator`` (pyrayhf_b200/library.py) fuses every stage into one pass and never
materialises the ``[n_freq x n_points]`` arrays.  The reference also exposes each stage as a public
function -- its tutorial notebook plots the output of ``regrid_to_nonuniform_grid`` -- so the same names
are offered here with the reference's signatures, units, return types and exceptions
(PyRayHF/library.py, "lib"):

    constants                  lib:40-72      (host constants, no kernel)
    den2freq                   lib:75-97
    find_X / find_Y            lib:120-158
    smooth_nonuniform_grid     lib:296-321
    regrid_to_nonuniform_grid  lib:324-438
    find_mu_mup                lib:161-256    (pyrayhf_b200.library.find_mu_mup)
    find_vh                    lib:259-293

numpy in -> numpy out (synchronous).  The arithmetic runs in the kernels of
``pyrayhf_b200/csrc/vfo_stages.cu`` behind the C ABI; there is no CPU fallback.
"""
import ctypes

import numpy as np

from pyrayhf_b200 import _cabi
from pyrayhf_b200.library import _f64, _mode_code, _raise_profile_status, find_mu_mup  # noqa: F401

_vp = ctypes.c_void_p


def constants():
    """``(cp, g_p, R_E, c_km_s)`` exactly as lib:60-72."""
    return 8.97866275, 2.799249247e10, 6371., 299_792.458


def _device():
    import torch
    return torch.device('cuda', torch.cuda.current_device())


def _to_dev(a, dev):
    import torch
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if not a.flags.writeable:                                 # e.g. a broadcast view: torch wants a writable buffer
        a = a.copy()
    return torch.from_numpy(a).to(dev)


def _stream(dev):
    import torch
    return _vp(torch.cuda.current_stream(dev).cuda_stream)


def _operand(a, shape, dev):
    """Device copy of one elementwise operand and its stride: 0 when it is a scalar broadcast."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 1 and int(np.prod(shape)) != 1:
        return _to_dev(a, dev), 0
    return _to_dev(np.broadcast_to(a, shape), dev), 1


def _result(t, shape):
    out = t.cpu().numpy().reshape(shape)
    return out[()] if out.ndim == 0 else out


def den2freq(density):
    """Plasma frequency [Hz] from density [m^-3]: ``sqrt(density) * cp`` (lib:75-97).

    Raises ``ValueError("Density must be non-negative")`` like lib:93-94.
    """
    import torch
    dev = _device()
    d = np.asarray(density, dtype=np.float64)
    td = _to_dev(d, dev)
    out = torch.empty_like(td)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_den2freq_f64(ctx.handle, _vp(td.data_ptr()), td.numel(), _vp(out.data_ptr()),
                                        _vp(flag.data_ptr()), _stream(dev)))
    if int(flag.item()):
        raise ValueError("Density must be non-negative")
    return _result(out, d.shape)


def find_X(n_e, f):
    """``X = (f_N / f)^2`` with the reference's rounding order (lib:120-137); operands broadcast."""
    import torch
    dev = _device()
    n_e = np.asarray(n_e, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    shape = np.broadcast_shapes(n_e.shape, f.shape)
    tn, sn = _operand(n_e, shape, dev)
    tf, sf = _operand(f, shape, dev)
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.float64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_find_x_f64(ctx.handle, _vp(tn.data_ptr()), sn, _vp(tf.data_ptr()), sf, n,
                                      _vp(out.data_ptr()), _vp(flag.data_ptr()), _stream(dev)))
    if int(flag.item()):
        raise ValueError("Density must be non-negative")                     # lib:94 through lib:136
    return _result(out, shape)


def find_Y(f, b):
    """``Y = g_p * b / f`` (lib:140-158); operands broadcast."""
    import torch
    dev = _device()
    f = np.asarray(f, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    shape = np.broadcast_shapes(f.shape, b.shape)
    tf, sf = _operand(f, shape, dev)
    tb, sb = _operand(b, shape, dev)
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.float64, device=dev)
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_find_y_f64(ctx.handle, _vp(tf.data_ptr()), sf, _vp(tb.data_ptr()), sb, n,
                                      _vp(out.data_ptr()), _stream(dev)))
    return _result(out, shape)


def smooth_nonuniform_grid(start, end, n_points, sharpness):
    """Smooth non-uniform grid, dense near ``end`` (lib:296-321)."""
    import torch
    n_points = int(n_points)
    if n_points < 0:
        raise ValueError("Number of samples, %d, must be non-negative." % n_points)   # np.linspace, lib:314
    dev = _device()
    out = torch.empty(n_points, dtype=torch.float64, device=dev)
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_smooth_grid_f64(ctx.handle, float(start), float(end), n_points, float(sharpness),
                                           _vp(out.data_ptr()), _stream(dev)))
    return out.cpu().numpy()


_REGRID_KEYS = ('freq', 'den', 'bmag', 'bpsi', 'dist', 'alt', 'crit_height', 'ind')


def regrid_to_nonuniform_grid(f, n_e, b, bpsi, aalt, mode='O', n_points=200, dh=1e-6, *, keys=None):
    """Regrid one profile to the stretched per-frequency grid (lib:324-438).

    ``f`` is in Hz, as in the reference.  Returns the reference's dict: ``'freq'``, ``'den'``, ``'bmag'``,
    ``'bpsi'``, ``'dist'``, ``'alt'``, ``'crit_height'`` (float64 ``[n_freq, n_points]``) and ``'ind'`` (integer
    ``[n_freq, n_points]``).  ``dh`` is accepted and ignored exactly as the reference ignores it (lib:378).
    ``keys`` (extension) restricts which entries are produced.
    """
    import torch
    code = _mode_code(mode)
    n_points = int(n_points)
    f = np.asarray(f, dtype=np.float64)
    if f.ndim > 1:
        raise ValueError("operands could not be broadcast together: f must be 0-d or 1-d")
    f = f.reshape(-1)
    d, bm, ps, al = (_f64(v).reshape(-1) for v in (n_e, b, bpsi, aalt))
    n_alt, n_freq = d.size, f.size
    if not (bm.size == ps.size == al.size == n_alt):
        raise ValueError("n_e, b, bpsi, aalt must have the same length")
    if n_alt == 0:
        raise ValueError("attempt to get argmax of an empty sequence")       # np.argmax, lib:371
    if n_freq == 0:
        raise ValueError("Cannot apply_along_axis when any iteration dimensions are 0")   # lib:403
    if n_points < 0:
        raise ValueError("Number of samples, %d, must be non-negative." % n_points)     # np.linspace, lib:314
    want = set(_REGRID_KEYS if keys is None else keys)
    unknown = want - set(_REGRID_KEYS)
    if unknown:
        raise KeyError("unknown regrid keys: %s" % sorted(unknown))
    dev = _device()
    tf, td, tb, tp, ta = (_to_dev(v, dev) for v in (f, d, bm, ps, al))
    hc = torch.empty(n_freq, dtype=torch.float64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    big = {k: (torch.empty((n_freq, n_points), dtype=torch.float64, device=dev)
               if (k in want and n_points > 0) else None)
           for k in ('alt', 'dist', 'den', 'bmag', 'bpsi')}
    ptr = lambda t: _vp(t.data_ptr()) if t is not None else None   # noqa: E731
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_regrid_f64(ctx.handle, ptr(tf), n_freq, ptr(td), ptr(tb), ptr(tp), ptr(ta), n_alt, code,
                                      max(n_points, 1), ptr(hc), ptr(big['alt']), ptr(big['dist']), ptr(big['den']),
                                      ptr(big['bmag']), ptr(big['bpsi']), ptr(st), _stream(dev)))
    status = int(st.item())
    if status:
        _raise_profile_status(status)
    out = {}
    for k in ('den', 'bmag', 'bpsi', 'dist', 'alt'):
        if k in want:
            out[k] = big[k].cpu().numpy() if big[k] is not None else np.empty((n_freq, 0))
    if 'freq' in want:
        out['freq'] = np.transpose(np.full((n_points, n_freq), f))           # lib:427
    if 'crit_height' in want:
        out['crit_height'] = np.transpose(np.broadcast_to(hc.cpu().numpy(), (n_points, n_freq)))   # lib:411-412
    if 'ind' in want:
        out['ind'] = np.full((n_freq, n_points), np.arange(0, n_points, 1))  # lib:418
    return {k: out[k] for k in _REGRID_KEYS if k in out}


def find_vh(X, Y, bpsi, dh, alt_min, mode, *, literal=False):
    """Virtual height of every row of ``[n_freq, n_points]`` arrays (lib:259-293)."""
    import torch
    X, Y, bpsi, dh = (np.asarray(v, dtype=np.float64) for v in (X, Y, bpsi, dh))
    shape = np.broadcast_shapes(X.shape, Y.shape, bpsi.shape, dh.shape)
    if len(shape) != 2:
        raise np.exceptions.AxisError("axis 1 is out of bounds for array of dimension %d" % len(shape))   # lib:288
    if mode not in ('O', 'X'):
        with np.errstate(all='ignore'):
            iso = Y.size > 0 and bool(np.nanmax(np.abs(Y)) < 1e-12)
        if not iso:
            raise ValueError("Mode must be O or X")                          # lib:226 (magnetised branch only)
        code = 0
    else:
        code = 0 if mode == 'O' else 1
    if shape[1] == 0 and shape[0] > 0:
        raise ValueError("zero-size array to reduction operation fmax which has no identity")   # lib:201
    dev = _device()
    tx, ty, tp, tdh = (_to_dev(np.broadcast_to(v, shape), dev) for v in (X, Y, bpsi, dh))
    vh = torch.empty(shape[0], dtype=torch.float64, device=dev)
    ctx = _cabi.context(dev.index)
    ctx.check(ctx.lib.prhf_find_vh_f64(ctx.handle, _vp(tx.data_ptr()), _vp(ty.data_ptr()), _vp(tp.data_ptr()),
                                       _vp(tdh.data_ptr()), shape[0], shape[1], float(alt_min), code,
                                       _cabi.FLAG_LITERAL if literal else 0, _vp(vh.data_ptr()), _stream(dev)))
    return vh.cpu().numpy()

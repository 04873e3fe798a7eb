"""Stratified Snell's-law ray tracers on the GPU, batched over rays.

Mirrors ``PyRayHF.library.trace_ray_cartesian_snells`` (PyRayHF/library.py:1096-1268) and
``trace_ray_spherical_snells`` (library.py:1460-1713): same arguments, units, dictionary keys and NaN
behaviour for one ray, plus ``trace_rays_snells_batched`` for a whole fan of (frequency, elevation) pairs over
one profile -- the reference needs one Python call (tens of milliseconds) per ray.  The arithmetic runs in
``pyrayhf_b200/csrc/vfo_snell.cu`` behind ``prhf_snell_f64``; there is no CPU fallback.
"""
import ctypes

import numpy as np

from pyrayhf_b200 import _cabi

_vp = ctypes.c_void_p
_KEYS_FAILED_SPHERICAL = ("x", "z", "group_path_km", "group_delay_sec", "x_midpoint", "z_midpoint",
                          "ground_range_km")
_KEYS = _KEYS_FAILED_SPHERICAL + ("x_apex_km", "z_apex_km")


def trace_rays_snells_batched(f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode='O', *, geometry='cartesian',
                              return_paths=False, dz_target_km=1.0, apex_boost=200.0, max_substeps=400, R_E=None,
                              literal=False):
    """Trace ``R`` rays through one stratified profile.

    ``f0_Hz`` and ``elevation_deg`` broadcast to a common 1-D shape ``[R]``.  Returns a dict of float64 arrays
    ``group_path_km, group_delay_sec, x_midpoint, z_midpoint, ground_range_km, x_apex_km, z_apex_km`` of shape
    ``[R]`` (NaN where the reference would return its all-NaN dict) and ``n_path`` (int32, points on each path,
    0 = no ray).  ``return_paths=True`` adds ``x`` and ``z`` of shape ``[R, 2 (n_alt + 1) + 1]``, NaN-padded.
    """
    import torch
    if geometry not in ('cartesian', 'spherical'):
        raise ValueError("geometry must be 'cartesian' or 'spherical'")
    f0, el = np.broadcast_arrays(np.asarray(f0_Hz, dtype=np.float64), np.asarray(elevation_deg, dtype=np.float64))
    # A (frequency x elevation) fan -- f0_Hz of shape [F, 1] against elevations of shape [E] or [1, E] -- goes through
    # the fan entry: the refractive-index field depends on the frequency only and is computed once per frequency there.
    fan = None
    if f0.ndim == 2 and f0.shape[0] > 0 and f0.shape[1] > 1 and f0.strides[1] == 0 and el.strides[0] == 0:
        fan = (np.ascontiguousarray(f0[:, 0]), np.ascontiguousarray(el[0, :]))
    f0 = np.ascontiguousarray(f0).reshape(-1)
    el = np.ascontiguousarray(el).reshape(-1)
    alt, ne, bb, ps = (np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for v in (alt_km, Ne, Babs, bpsi))
    n_alt, n_rays = alt.size, f0.size
    if not (ne.size == bb.size == ps.size == n_alt) or n_alt == 0:
        raise ValueError("alt_km, Ne, Babs, bpsi must be non-empty and have the same length")
    if np.any(ne < 0):
        raise ValueError("Density must be non-negative")                     # library.py:94 through find_X
    if mode == 'O':
        code = 0
    elif mode == 'X':
        code = 1
    else:
        # find_mu_mup only looks at the mode on its magnetised branch (library.py:201, 221-226)
        with np.errstate(all='ignore'):
            ymax = np.nanmax(np.abs(bb)) * 2.799249247e10 / np.min(np.abs(f0)) if n_rays else 0.0
        if not (ymax < 1e-12):
            raise ValueError("Mode must be O or X")
        code = 0
    r_e = 6371.0 if R_E is None else float(R_E)
    dev = torch.device('cuda', torch.cuda.current_device())
    t_f, t_e, t_a, t_n, t_b, t_p = (torch.from_numpy(v).to(dev) for v in (f0, el, alt, ne, bb, ps))
    scal = torch.empty((n_rays, 5), dtype=torch.float64, device=dev)
    n_path = torch.zeros(n_rays, dtype=torch.int32, device=dev)
    stride = 2 * (n_alt + 1) + 1
    xs = torch.empty((n_rays, stride), dtype=torch.float64, device=dev) if return_paths else None
    zs = torch.empty((n_rays, stride), dtype=torch.float64, device=dev) if return_paths else None
    ptr = lambda t: _vp(t.data_ptr()) if t is not None else None             # noqa: E731
    ctx = _cabi.context(dev.index)
    tail = (ptr(t_a), ptr(t_n), ptr(t_b), ptr(t_p), n_alt, code, 1 if geometry == 'spherical' else 0,
            _cabi.FLAG_LITERAL if literal else 0, float(dz_target_km), float(apex_boost), int(max_substeps), r_e,
            ptr(scal), ptr(xs), ptr(zs), stride, ptr(n_path), _vp(torch.cuda.current_stream(dev).cuda_stream))
    if fan is not None:
        t_ff, t_fe = (torch.from_numpy(v).to(dev) for v in fan)
        ctx.check(ctx.lib.prhf_snell_fan_f64(ctx.handle, ptr(t_ff), fan[0].size, ptr(t_fe), fan[1].size, *tail))
    else:
        ctx.check(ctx.lib.prhf_snell_f64(ctx.handle, ptr(t_f), ptr(t_e), n_rays, *tail))
    s = scal.cpu().numpy()
    out = {"group_path_km": s[:, 0].copy(), "group_delay_sec": s[:, 1].copy(), "x_midpoint": s[:, 2].copy(),
           "z_midpoint": s[:, 3].copy(), "ground_range_km": s[:, 4].copy(), "x_apex_km": s[:, 2].copy(),
           "z_apex_km": s[:, 3].copy(), "n_path": n_path.cpu().numpy()}
    if return_paths:
        out["x"] = xs.cpu().numpy()
        out["z"] = zs.cpu().numpy()
    return out


def _single(geometry, f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode, **kw):
    r = trace_rays_snells_batched(np.array([float(f0_Hz)]), np.array([float(elevation_deg)]), alt_km, Ne, Babs, bpsi,
                                  mode, geometry=geometry, return_paths=True, **kw)
    n = int(r["n_path"][0])
    if n == 0:
        # the reference's early exits: every key NaN; the spherical ones omit the apex keys (library.py:1578-1585)
        return {k: np.nan for k in (_KEYS if geometry == 'cartesian' else _KEYS_FAILED_SPHERICAL)}
    out = {"x": r["x"][0, :n].copy(), "z": r["z"][0, :n].copy()}
    for k in _KEYS[2:]:
        out[k] = float(r[k][0])
    return out


def trace_ray_cartesian_snells(f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode):
    """One flat-Earth ray; drop-in for library.py:1096-1268 (same dict)."""
    return _single('cartesian', f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode)


def trace_ray_spherical_snells(f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode="O", *, dz_target_km=1.0,
                               apex_boost=200.0, max_substeps=400, R_E=None):
    """One spherical-Earth ray; drop-in for library.py:1460-1713 (same dict, same keyword controls)."""
    return _single('spherical', f0_Hz, elevation_deg, alt_km, Ne, Babs, bpsi, mode, dz_target_km=dz_target_km,
                   apex_boost=apex_boost, max_substeps=max_substeps, R_E=R_E)

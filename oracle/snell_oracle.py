"""TEST INFRASTRUCTURE ONLY -- numpy restatement of PyRayHF's stratified Snell's-law ray tracers.

Restates ``trace_ray_cartesian_snells`` (PyRayHF/library.py:1096-1268, helpers ``tan_from_mu_scalar``
library.py:1034-1062 and ``find_turning_point`` library.py:1065-1093) and ``trace_ray_spherical_snells``
(library.py:1460-1713) as ONE routine with a geometry switch, vectorised over the profile levels (the reference
loops in Python).  The refractive-index field comes from ``oracle.vfo_oracle`` (find_X / find_Y / find_mu_mup
restatements).

Parity status: PINNED against the live reference (tests/test_oracle_vs_reference.py::test_snell_*) and against
tests/golden/snell.npz captured from it (tests/make_golden_snell.py).

One output pair is rounding-dependent IN THE REFERENCE: the path is mirror-symmetric, so half of the path
length coincides (to rounding) with the cumulative length at the apex, and ``np.searchsorted`` lands on the apex
or on the node just below it depending on the last bit.  ``trace`` reports both candidates
(``mid_candidates``) so that the tests can accept either.

Only ``tests/`` may import this module; the product path (``pyrayhf_b200``) never does.
"""
import numpy as np

from oracle import vfo_oracle

C_KM_S = 299_792.458          # library.py:70
R_EARTH_KM = 6371.0           # library.py:67
KEYS = ("x", "z", "group_path_km", "group_delay_sec", "x_midpoint", "z_midpoint", "ground_range_km",
        "x_apex_km", "z_apex_km")


def _field(f0_hz, alt, ne, babs, bpsi, mode):
    """Ground level inserted, then mu / mu' with non-positive and non-finite values masked.

    library.py:1181-1200 (cartesian) == library.py:1560-1575 (spherical).
    """
    alt, ne, babs, bpsi = (np.asarray(v, dtype=float) for v in (alt, ne, babs, bpsi))
    if alt[0] > 0.0:
        ne = np.concatenate(([np.interp(0.0, alt, ne)], ne))
        babs = np.concatenate(([np.interp(0.0, alt, babs)], babs))
        bpsi = np.concatenate(([np.interp(0.0, alt, bpsi)], bpsi))
        alt = np.concatenate(([0.0], alt))
    x = vfo_oracle.plasma_ratio_x(ne, f0_hz)
    y = vfo_oracle.gyro_ratio_y(f0_hz, babs)
    with np.errstate(all='ignore'):
        mu, mup = vfo_oracle.appleton_hartree(x, y, bpsi, mode)
        mu = np.where(np.isfinite(mu) & (mu > 0.0), mu, np.nan)
        mup = np.where(np.isfinite(mup) & (mup > 0.0), mup, np.nan)
    return alt, mu, mup


def _failed(geometry):
    keys = KEYS if geometry == 'cartesian' else KEYS[:7]     # the spherical early exits omit the apex keys
    return {k: np.nan for k in keys}


def trace(f0_hz, elevation_deg, alt_km, ne, babs, bpsi, mode, geometry='cartesian', *, dz_target_km=1.0,
          apex_boost=200.0, max_substeps=400, r_e=None):
    """One ray.  Returns the reference's dict plus ``mid_candidates`` (indices of the two admissible midpoints)."""
    sph = geometry == 'spherical'
    r_e = R_EARTH_KM if r_e is None else r_e
    alt, mu, mup = _field(f0_hz, alt_km, ne, babs, bpsi, mode)
    s0 = np.sin(np.radians(90.0 - elevation_deg))
    if not np.isfinite(mu[0]) or (not sph and not np.isfinite(s0)):
        return _failed(geometry)
    radius = (r_e + alt) if sph else np.ones_like(alt)
    p = mu[0] * radius[0] * s0 if sph else mu[0] * s0                        # lib:1211 / lib:1588
    ok = np.isfinite(mu)
    zv, muv, rv = alt[ok], mu[ok], radius[ok]
    if zv.size < 2:
        return _failed(geometry)
    q = muv * rv if sph else muv                                             # the quantity that crosses p
    hit = np.flatnonzero((q[:-1] >= p) & (q[1:] <= p))
    if hit.size == 0:
        return _failed(geometry)
    i0 = int(hit[0])
    if q[i0] == q[i0 + 1]:
        t = 0.0
    else:
        t = (q[i0] - p) / (q[i0] - q[i0 + 1])
    if sph:
        t = float(np.clip(t, 0.0, 1.0))                                      # lib:1622
    z_turn = zv[i0] + t * (zv[i0 + 1] - zv[i0]) if (sph or q[i0] != q[i0 + 1]) else zv[i0]
    if sph:
        z_up = np.concatenate((zv[:i0 + 1], [z_turn]))                       # lib:1626
        mu_up = np.concatenate((muv[:i0 + 1], [p / (r_e + z_turn)]))
        r_up = r_e + z_up
        phi = np.zeros_like(z_up)
        for k in range(z_up.size - 1):
            dz = z_up[k + 1] - z_up[k]
            if dz <= 0:
                continue                                                     # leaves phi[k + 1] == 0 (lib:1647)
            qa, qb = mu_up[k] * r_up[k], mu_up[k + 1] * r_up[k + 1]
            n = max(1, int(np.ceil(abs(dz) / dz_target_km)))
            sharp = 1.0 / min(max(qa - p, 1e-12), max(qb - p, 1e-12))
            n = int(min(max_substeps, n * (1.0 + apex_boost * sharp)))
            j = np.arange(n)
            tm = 0.5 * (j / n + (j + 1) / n)
            rm = r_e + (z_up[k] + tm * dz)
            qm = (mu_up[k] + (mu_up[k + 1] - mu_up[k]) * tm) * rm
            qm = np.where(qm <= p, p + 1e-8, qm)
            fm = p / (rm * np.sqrt(np.maximum(qm * qm - p * p, 1e-16)))
            acc = 0.0
            for term in fm * (dz / n):                                       # sequential, as lib:1661-1673
                acc += term
            phi[k + 1] = phi[k] + acc
        phi_full = np.concatenate((phi, (2.0 * phi[-1] - phi[::-1])[1:]))
        z_full = np.concatenate((z_up, z_up[::-1][1:]))
        x_full = r_e * phi_full
        r_mid = r_e + 0.5 * (z_full[:-1] + z_full[1:])
        ds = np.hypot(r_mid * np.diff(phi_full), np.diff(z_full))
    else:
        i_turn = int(np.searchsorted(zv, z_turn))                            # lib:1228
        z_up = np.concatenate((zv[:i_turn], [z_turn]))
        mu_up = np.concatenate((muv[:i_turn], [p]))
        x_up = np.zeros_like(z_up)
        if z_up.size > 1:
            mid = 0.5 * (mu_up[:-1] + mu_up[1:])
            mid[-1] = max(mid[-1], p + 1e-8)
            tan_mid = p / np.sqrt(np.maximum(mid * mid - p * p, 1e-10))      # lib:1056-1061
            x_up[1:] = np.cumsum(np.diff(z_up) * tan_mid)
        x_full = np.concatenate((x_up, (2.0 * x_up[-1] - x_up[::-1])[1:]))
        z_full = np.concatenate((z_up, z_up[::-1][1:]))
        ds = np.hypot(np.diff(x_full), np.diff(z_full))
    path = float(np.nansum(ds))
    mup_path = np.interp(z_full, alt, mup)
    delay = float(np.nansum((0.5 * (mup_path[1:] + mup_path[:-1]) / C_KM_S) * ds))
    n_up = z_up.size
    if path > 0:
        mid_idx = int(np.searchsorted(np.cumsum(ds), 0.5 * path))
        xm, zm = float(x_full[mid_idx]), float(z_full[mid_idx])
    else:
        mid_idx, xm, zm = -1, np.nan, np.nan
    ground = float(x_full[-1]) if np.isclose(z_full[-1], 0.0, atol=1e-3) else np.nan
    return {"x": x_full, "z": z_full, "group_path_km": path, "group_delay_sec": delay, "x_midpoint": xm,
            "z_midpoint": zm, "ground_range_km": ground, "x_apex_km": xm, "z_apex_km": zm,
            "mid_index": mid_idx, "mid_candidates": (max(n_up - 2, 0), n_up - 1)}

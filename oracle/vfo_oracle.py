"""TEST INFRASTRUCTURE ONLY -- numpy restatement of PyRayHF's vertical forward operator.

This is the CPU *oracle* for the hot path.  It restates, stage by stage, what
``PyRayHF/library.py`` computes on the path ``vertical_forward_operator``
(library.py:459-509) -> ``regrid_to_nonuniform_grid`` (library.py:324-438) ->
``find_X``/``find_Y`` (library.py:120-158) -> ``find_vh`` (library.py:259-293)
-> ``find_mu_mup`` (library.py:161-256).  It deliberately keeps the reference's
whole-array ``[n_freq x n_points]`` evaluation style and the same numpy
primitives (``np.interp``, ``np.maximum.accumulate``, ``np.nansum``, ``**4``),
so that on one machine / one numpy build it is bit-identical to the reference
and has the same CPU cost profile (it doubles as the ``cpu_baseline`` "port").

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` checks it
bit-for-bit against the live reference (when ``/root/reference`` is mounted)
and ``tests/test_oracle_golden.py`` against the committed fixtures in
``tests/golden`` that ``tests/make_golden.py`` generated from the live
reference, including the reference's own known-answer tests
(tests/test_core.py:137-152, 223-236, 239-276).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``pyrayhf_b200``) never does.
"""
import numpy as np

# library.py:61 and library.py:64
CP_HZ_PER_SQRT_M3 = 8.97866275
GYRO_HZ_PER_T = 2.799249247e10
# library.py:363 and library.py:378 (both hard-coded inside the path)
SHARPNESS = 10.0
BACKOFF_KM = 1e-6
# library.py:163
Y_TOL = 1e-12


def stretch_multiplier(n_points, sharpness=SHARPNESS):
    """Stretched-grid multiplier m_i in [0, 1], dense near 1.

    library.py:314-320 called with (start, end) = (0, 1) from library.py:361-364.
    """
    u = np.linspace(0.0, 1.0, n_points)
    factor = (np.exp(sharpness * (1.0 - u)) - 1.0) / (np.exp(sharpness) - 1.0)
    return 1.0 - (0.0 + (1.0 - 0.0) * factor)


def smooth_grid(start, end, n_points, sharpness):
    """General stretched grid.  library.py:296-321."""
    u = np.linspace(0.0, 1.0, n_points)
    factor = (np.exp(sharpness * (1.0 - u)) - 1.0) / (np.exp(sharpness) - 1.0)
    return 1. - (start + (end - start) * factor)


def den2freq(den):
    """sqrt(density) * cp with the negative-density check.  library.py:75-97."""
    if np.any(np.asarray(den) < 0):
        raise ValueError("Density must be non-negative")
    return np.sqrt(den) * CP_HZ_PER_SQRT_M3


def plasma_ratio_x(den, f_hz):
    """X = (sqrt(n) * cp)**2 / f**2 with the reference's rounding order.

    library.py:93-96 (negative density check, sqrt then scale) and library.py:136.
    """
    if np.any(np.asarray(den) < 0):
        raise ValueError("Density must be non-negative")
    return (np.sqrt(den) * CP_HZ_PER_SQRT_M3) ** 2 / f_hz ** 2


def gyro_ratio_y(f_hz, bmag):
    """Y = g_p * B / f, multiply first.  library.py:157."""
    return GYRO_HZ_PER_T * bmag / f_hz


def truncate_below_peak(den, bmag, bpsi, alt):
    """Keep samples strictly below the density peak.  library.py:371-375."""
    k = int(np.argmax(den))
    return den[:k], bmag[:k], bpsi[:k], alt[:k]


def reflection_heights(f_hz, den_t, bmag_t, alt_t, mode):
    """Per-frequency reflection height (already backed off by 1e-6 km).

    library.py:380-407.  Returns (h_c[F] with NaN on dead rows, valid[F]).
    """
    n_freq, n_alt = f_hz.size, alt_t.size
    den2 = np.broadcast_to(den_t, (n_freq, n_alt))
    b2 = np.broadcast_to(bmag_t, (n_freq, n_alt))
    f2 = np.broadcast_to(f_hz, (n_alt, n_freq)).T
    x = plasma_ratio_x(den2, f2)
    y = gyro_ratio_y(f2, b2)
    if mode == 'O':
        crit = np.maximum.accumulate(x, axis=1)
    elif mode == 'X':
        crit = np.maximum.accumulate(x + y, axis=1)
    else:
        raise ValueError("mode must be 'O' or 'X'")
    valid = crit[:, -1] >= 1.0          # IndexError when n_alt == 0, as library.py:399
    h = np.empty(n_freq)
    for r in range(n_freq):             # library.py:403-404 (apply_along_axis)
        h[r] = np.interp(1.0, crit[r], alt_t)
    return np.where(valid, h - BACKOFF_KM, np.nan), valid


def regrid(f_hz, den, bmag, bpsi, alt, mode, n_points):
    """Stretched per-frequency altitude grid and the profile sampled on it.

    library.py:361-438.  Returns a dict with h, dh, den, bmag, bpsi ([F x N]),
    h_c[F], valid[F], multiplier[N].
    """
    m = stretch_multiplier(n_points)
    den_t, bmag_t, bpsi_t, alt_t = truncate_below_peak(den, bmag, bpsi, alt)
    h_c, valid = reflection_heights(f_hz, den_t, bmag_t, alt_t, mode)
    h = m[None, :] * (h_c[:, None] - alt_t[0]) + alt_t[0]           # library.py:413
    dh = np.concatenate((np.diff(h, axis=1),
                         np.full((f_hz.size, 1), BACKOFF_KM)), axis=1)  # library.py:415-416
    flat = h.reshape(-1)
    out = {'h': h, 'dh': dh, 'h_c': h_c, 'valid': valid, 'multiplier': m}
    for key, tab in (('den', den_t), ('bmag', bmag_t), ('bpsi', bpsi_t)):
        out[key] = np.interp(flat, alt_t, tab).reshape(h.shape)     # library.py:424-426
    return out


def regrid_dict(f_hz, den, bmag, bpsi, alt, mode, n_points):
    """The reference's regridded dict (same keys and shapes).  library.py:418-438."""
    g = regrid(f_hz, den, bmag, bpsi, alt, mode, n_points)
    n_freq = f_hz.size
    return {'freq': np.transpose(np.full((n_points, n_freq), f_hz)),
            'den': g['den'], 'bmag': g['bmag'], 'bpsi': g['bpsi'], 'dist': g['dh'], 'alt': g['h'],
            'crit_height': np.transpose(np.broadcast_to(g['h_c'], (n_points, n_freq))),
            'ind': np.full((n_freq, n_points), np.arange(0, n_points, 1))}


def find_vh_rows(x, y, psi_deg, dh, alt_min, mode):
    """Row sums of mu' * dh.  library.py:259-293."""
    _, mup = appleton_hartree(x, y, psi_deg, mode)
    s = np.nansum(mup * dh, axis=1)
    s[s == 0] = np.nan
    return s + alt_min


def appleton_hartree(x, y, psi_deg, mode):
    """Phase index mu and group index mu' (Appleton-Hartree, collisionless).

    library.py:194-256, including the whole-array unmagnetised switch at
    library.py:201 and the NaN masks at library.py:233 and library.py:238.
    """
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    psi_deg = np.asarray(psi_deg, dtype=float)
    if np.nanmax(np.abs(y)) < Y_TOL:                                 # library.py:201-207
        mu2 = 1.0 - x
        mu = np.where(mu2 > 0.0, np.sqrt(mu2), np.nan)
        mup = np.where(np.isfinite(mu) & (mu > 0.0), 1.0 / mu, np.nan)
        return mu, mup
    if mode == 'O':
        sgn = 1.0
    elif mode == 'X':
        sgn = -1.0
    else:
        raise ValueError("Mode must be O or X")
    sin_p = np.sin(np.deg2rad(psi_deg))
    cos_p = np.cos(np.deg2rad(psi_deg))
    yt = y * sin_p
    yl = y * cos_p
    xm1 = 1.0 - x
    beta = np.sqrt(0.25 * yt ** 4 + yl ** 2 * xm1 ** 2)              # library.py:217-218
    d = xm1 - 0.5 * yt ** 2 + sgn * beta                             # library.py:229
    u = 1.0 - x * xm1 / d                                            # library.py:232
    u[u < 0] = np.nan
    mu = np.sqrt(u)
    mu[np.where(mu < 0.0)] = 0.0
    mu[np.where(mu > 1.0)] = np.nan
    dbeta_dx = -yl ** 2 * xm1 / beta                                 # library.py:241
    dd_dx = -1.0 + sgn * dbeta_dx
    dalpha_dy = (yt ** 3 * sin_p) + (2.0 * yl * xm1 ** 2 * cos_p)    # library.py:244-245
    dbeta_dy = 0.5 * dalpha_dy / beta
    dd_dy = -yt * sin_p + sgn * dbeta_dy                             # library.py:247
    dmu_dy = (x * xm1 * dd_dy) / (2.0 * mu * d ** 2)                 # library.py:250
    dmu_dx = (1.0 / (2.0 * mu * d)) * (2.0 * x - 1.0 + x * xm1 / d * dd_dx)
    mup = mu - (2.0 * x * dmu_dx + y * dmu_dy)                       # library.py:254
    return mu, mup


def vertical_forward_operator(freq, den, bmag, bpsi, alt, mode='O',
                              n_points=200, stages=False):
    """Virtual height [km] per sounding frequency [MHz].  library.py:459-509.

    ``stages=True`` additionally returns the intermediate arrays for
    stage-by-stage diffing against the CUDA path.
    """
    f_hz = freq * 1e6                                                # library.py:491
    g = regrid(f_hz, den, bmag, bpsi, alt, mode, n_points)
    f2 = np.broadcast_to(f_hz[:, None], g['h'].shape)
    x = plasma_ratio_x(g['den'], f2)                                 # library.py:500
    y = gyro_ratio_y(f2, g['bmag'])                                  # library.py:503
    mu, mup = appleton_hartree(x, y, g['bpsi'], mode)                # library.py:285
    s = np.nansum(mup * g['dh'], axis=1)                             # library.py:288
    s[s == 0] = np.nan                                               # library.py:290
    vh = s + np.min(alt)                                             # library.py:292, 507
    if stages:
        g.update(X=x, Y=y, mu=mu, mup=mup)
        return vh, g
    return vh


def vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, mode='O',
                                      n_points=200):
    """Row p of the result == the reference called on profile p alone.

    ``freq`` is [F] or [P, F]; ``alt`` is [A] or [P, A]; den/bmag/bpsi [P, A].
    Per-profile errors of the reference (library.py:94, library.py:399) are
    re-raised; use ``profile_status`` to pre-screen.
    """
    den = np.asarray(den)
    n_prof = den.shape[0]
    freq = np.asarray(freq)
    alt = np.asarray(alt)
    out = np.empty((n_prof, freq.shape[-1]))
    for p in range(n_prof):
        fp = freq[p] if freq.ndim == 2 else freq
        ap = alt[p] if alt.ndim == 2 else alt
        out[p] = vertical_forward_operator(fp, den[p], bmag[p], bpsi[p], ap,
                                           mode, n_points)
    return out


def residual_from_model(vh_obs, vh_model):
    """Arithmetic tail of residual_VH: NaN fill and difference.  library.py:660-668.

    ``vh_model`` is one modelled curve [F] (as in the reference) or a batch [P, F] (row-wise).
    """
    vh_model = np.array(vh_model, dtype=float, copy=True)
    if vh_model.ndim == 1:
        vh_model[np.isnan(vh_model)] = np.maximum(np.nanmean(np.abs(vh_model)), 100)   # library.py:664-665
        return (vh_obs - vh_model).ravel()                                               # library.py:668
    return np.stack([residual_from_model(vh_obs, row) for row in vh_model])


def profile_status(den):
    """0 ok / 1 negative density below the peak / 2 peak at index 0.

    Mirrors where the reference raises: ValueError at library.py:94 (via
    library.py:384) is hit before the IndexError at library.py:399 only if the
    truncated profile is non-empty, so an index-0 peak reports 2.
    """
    k = int(np.argmax(den))
    if k == 0:
        return 2
    if np.any(den[:k] < 0):
        return 1
    return 0

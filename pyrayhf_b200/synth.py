"""Synthetic Chapman-layer ionosphere + centred-dipole field inputs.

PyIRI / IGRF (what ``generate_input_1D`` uses, library.py:2390-2694) are not
available offline, so benchmarks and parity tests use this deterministic
generator instead (definition: SURVEY.md section 8d "Common inputs").  It is
input synthesis on the host, outside the timed hot path.
"""
import numpy as np

CP_HZ_PER_SQRT_M3 = 8.97866275      # library.py:61
R_EARTH_KM = 6371.0                 # library.py:67
B0_TESLA = 3.12e-5


def default_alt():
    """80 ... 699 km in 1 km steps, as the tutorial inputs (A = 620)."""
    return np.arange(80.0, 700.0, 1.0)


def default_freq():
    """0.1 ... 17.4 MHz in 0.1 MHz steps (F = 174), README example."""
    return np.arange(0.1, 17.5, 0.1)


def chapman(alt, nm, hm, scale_h):
    z = (alt - hm) / scale_h
    return nm * np.exp(0.5 * (1.0 - z - np.exp(-z)))


def layer_parameters(lat_deg, lon_deg):
    """Deterministic (foF2, hmF2, H, foE) for geographic points (broadcasts)."""
    lat = np.deg2rad(np.asarray(lat_deg, dtype=float))
    lon = np.deg2rad(np.asarray(lon_deg, dtype=float))
    c = np.cos(lat)
    d = 0.55 + 0.45 * np.cos(lon)
    fof2 = 3.0 + 10.0 * c * c * d
    hmf2 = 250.0 + 100.0 * c * c * d
    scale_h = 40.0 + 20.0 * d
    foe = 0.5 + 3.0 * np.sqrt(np.clip(c * d, 0.0, None))
    return fof2, hmf2, scale_h, foe


def profiles_from_parameters(fof2, hmf2, scale_h, foe, lat_deg, alt=None):
    """[P, A] den / bmag / bpsi for arrays of layer parameters (length P)."""
    alt = default_alt() if alt is None else np.asarray(alt, dtype=float)
    fof2, hmf2, scale_h, foe, lat_deg = (np.atleast_1d(np.asarray(v, dtype=float))
                                         for v in (fof2, hmf2, scale_h, foe, lat_deg))
    nmf2 = (fof2 * 1e6 / CP_HZ_PER_SQRT_M3) ** 2
    nme = (foe * 1e6 / CP_HZ_PER_SQRT_M3) ** 2
    a = alt[None, :]
    den = (chapman(a, nmf2[:, None], hmf2[:, None], scale_h[:, None])
           + chapman(a, nme[:, None], 110.0, 8.0))
    lat = np.deg2rad(lat_deg)[:, None]
    bmag = (B0_TESLA * (R_EARTH_KM / (R_EARTH_KM + a)) ** 3
            * np.sqrt(1.0 + 3.0 * np.sin(lat) ** 2))
    incl = np.rad2deg(np.arctan2(2.0 * np.sin(lat), np.cos(lat)))
    bpsi = np.broadcast_to(90.0 - np.abs(incl), den.shape).copy()
    return den, np.ascontiguousarray(bmag), bpsi


def profiles_from_parameters_device(fof2, hmf2, scale_h, foe, lat_deg, alt=None, device=None):
    """``profiles_from_parameters`` evaluated by the CUDA generator kernel: float64 CUDA tensors
    ``(den, bmag, bpsi)`` of shape ``[P, A]`` that never touch the host (config 4 builds its 8 192-profile
    chunks this way, so the ensemble needs no host-to-device traffic beyond 40 bytes per profile)."""
    import ctypes
    import torch
    from pyrayhf_b200 import _cabi
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    alt = default_alt() if alt is None else np.asarray(alt, dtype=float)
    cols = [np.atleast_1d(np.asarray(v, dtype=float)) for v in (fof2, hmf2, scale_h, foe, lat_deg)]
    params = torch.from_numpy(np.ascontiguousarray(np.stack(np.broadcast_arrays(*cols), axis=1))).to(dev)
    t_alt = torch.from_numpy(np.ascontiguousarray(alt)).to(dev)
    n_prof, n_alt = params.shape[0], t_alt.numel()
    den, bmag, bpsi = (torch.empty((n_prof, n_alt), dtype=torch.float64, device=dev) for _ in range(3))
    ctx = _cabi.context(dev.index)
    vp = ctypes.c_void_p
    ctx.check(ctx.lib.prhf_synth_profiles_f64(ctx.handle, vp(params.data_ptr()), n_prof, vp(t_alt.data_ptr()), n_alt,
                                              vp(den.data_ptr()), vp(bmag.data_ptr()), vp(bpsi.data_ptr()),
                                              vp(torch.cuda.current_stream(dev).cuda_stream)))
    return den, bmag, bpsi


def ensemble_member_parameters(lat_deg, lon_deg, member):
    """Layer parameters of one config-4 ensemble member (the perturbation of ``ensemble_member``)."""
    lat_deg = np.atleast_1d(np.asarray(lat_deg, dtype=float))
    fof2, hmf2, scale_h, foe = layer_parameters(lat_deg, lon_deg)
    xi = np.random.default_rng(1000 + int(member)).standard_normal((3, lat_deg.size))
    return fof2 * np.exp(0.05 * xi[0]), hmf2 + 10.0 * xi[1], scale_h * np.exp(0.05 * xi[2]), foe, lat_deg


def profiles_at(lat_deg, lon_deg, alt=None):
    """[P, A] den / bmag / bpsi for geographic points (1-D arrays of length P)."""
    lat_deg = np.atleast_1d(np.asarray(lat_deg, dtype=float))
    lon_deg = np.atleast_1d(np.asarray(lon_deg, dtype=float))
    fof2, hmf2, scale_h, foe = layer_parameters(lat_deg, lon_deg)
    return profiles_from_parameters(fof2, hmf2, scale_h, foe, lat_deg, alt)


def single_day_profile(alt=None):
    """The config-1/2 synthetic day profile (lat 4.5, lon -150)."""
    alt = default_alt() if alt is None else alt
    den, bmag, bpsi = profiles_at([4.5], [-150.0], alt)
    return den[0], bmag[0], bpsi[0], alt


def global_grid_points():
    """Config 3: the 181 x 361 one-degree lat/lon grid, P = 65 341."""
    lat, lon = np.meshgrid(np.arange(-90.0, 91.0, 1.0), np.arange(-180.0, 181.0, 1.0),
                           indexing="ij")
    return lat.reshape(-1), lon.reshape(-1)


def grid_subset(n_profiles, seed=20260101):
    """First ``n_profiles`` of the seeded shuffle of the global grid (config 4 base)."""
    lat, lon = global_grid_points()
    order = np.random.default_rng(seed).permutation(lat.size)[:n_profiles]
    return lat[order], lon[order]


def ensemble_member(lat_deg, lon_deg, member, alt=None):
    """Config 4: perturbed copy of the base profiles for one ensemble member."""
    # NmF2 * exp(0.10 xi1)  <=>  foF2 * exp(0.05 xi1)
    return profiles_from_parameters(*ensemble_member_parameters(lat_deg, lon_deg, member), alt)


def bench_day_profile(alt=None, rank=0):
    """Benchmark profile for BASELINE configs[1]: day side of the parameter map (lat 4.5, lon 0).

    foF2 = 12.9 MHz, so 129 of the 174 sounding frequencies reflect (the real tutorial Day profile:
    133 of 174).  SURVEY.md 8d's (4.5, -150) point sits on the night side of its own map
    (foF2 = 4.6 MHz, 41 of 174 rows reflect) and would flatter a virtual-heights/s figure, because
    rows that never reflect cost the GPU almost nothing.  ``rank`` shifts the longitude by one degree
    per rank so that every GPU of a multi-GPU run owns a different profile.
    """
    alt = default_alt() if alt is None else alt
    den, bmag, bpsi = profiles_at([4.5], [0.0 + float(rank)], alt)
    return den[0], bmag[0], bpsi[0], alt

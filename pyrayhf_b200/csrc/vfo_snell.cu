// Stratified Snell's-law ray tracers, batched over rays (frequency, launch elevation) -- sm_100a, FP64.
//
// Reference behaviour being replaced (PyRayHF/library.py, "lib"): trace_ray_cartesian_snells (lib:1096-1268, with
// tan_from_mu_scalar lib:1034-1062 and find_turning_point lib:1065-1093) and trace_ray_spherical_snells
// (lib:1460-1713).  The reference traces ONE ray per call with Python loops over the profile levels (and, in the
// spherical case, up to 400 midpoint sub-steps per level); here one warp owns one ray and a launch covers the
// whole (frequency x elevation) fan over one shared profile.
//
// Per ray (all in shared memory, nothing but the results goes to HBM):
//   A  ground level inserted (lib:1169-1178), X, Y (lib:1181-1182), whole-array unmagnetised switch (lib:201),
//      mu / mu' per level with non-positive / non-finite values masked (lib:1183-1185)
//   B  levels with finite mu compacted (lib:1214-1215)
//   C  first crossing of the Snell invariant, linear turning altitude (lib:1080-1093 / lib:1603-1623)
//   D  horizontal coordinate of the up-leg: midpoint tan(theta) per level (lib:1233-1241) or the adaptive
//      midpoint rule on d(phi)/dz (lib:1633-1675); the running sum is sequential, in the reference's order
//   E  mirrored down-leg, segment lengths, group path and group delay (lib:1243-1256 / lib:1677-1695),
//      midpoint by the reference's searchsorted on the cumulative length (lib:1258-1264)
// Arithmetic that the reference's result is sensitive to (the near-apex (mu r)^2 - p^2 and mu^2 - p^2
// differences) is written with explicit round-to-nearest intrinsics so that no FMA contraction changes it.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "vfo_device.cuh"
#include "vfo_kernels.h"

namespace prhf {

namespace {

constexpr double kCkmS = 299792.458;     // lib:70

struct WarpRay {
  double* muv;    // [n]   compacted mu, later the up-leg segment lengths
  double* mup;    // [n]   mu' on the full grid (NaN-masked)
  double* xu;     // [n+1] mu on the full grid (phase A), then x_up / phi_up
  int* vidx;      // [n]   full-grid index of every compacted level
};

// Phase A for one sounding frequency: mu and mu' at the n levels (ground level inserted), NaN-masked as lib:1184-1185,
// written by the lanes of one warp to mu_out[k], mup_out[k].
template <int MODE, bool LITERAL>
__device__ __forceinline__ void field_levels(const SnellParams& p, double f0, int n, int ins, int lane, double* mu_out,
                                             double* mup_out) {
  const unsigned full = 0xffffffffu;
  double ymax = -1.0;                                        // nanmax |Y| (lib:201); -1 = no non-NaN value
  for (int k = lane; k < n; k += 32) {
    const int src = max(k - ins, 0);                         // np.interp(0, alt, v) clamps to v[0] (lib:1171-1173)
    const double y = fabs(y_literal(p.babs[src], f0));
    if (y == y) ymax = fmax(ymax, y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ymax = fmax(ymax, __shfl_xor_sync(full, ymax, o));
  const bool iso = (ymax >= 0.0) && (ymax < kYTol);
  for (int k = lane; k < n; k += 32) {
    const int src = max(k - ins, 0);
    const double X = x_literal(p.ne[src], f0);
    double mu, mup;
    if (iso) {
      mup = iso_mup(X, &mu);
    } else if (LITERAL) {
      mup = ah_literal<MODE>(X, y_literal(p.babs[src], f0), p.bpsi[src], &mu);
    } else {
      double sn, cs;
      sincos(p.bpsi[src] * kDeg2Rad, &sn, &cs);
      mup = ah_fast<MODE>(X, y_literal(p.babs[src], f0), sn, cs, &mu);
    }
    mu_out[k] = (isfinite(mu) && mu > 0.0) ? mu : CUDART_NAN;          // lib:1184
    mup_out[k] = (isfinite(mup) && mup > 0.0) ? mup : CUDART_NAN;      // lib:1185
  }
}

// Fan entry: the refractive-index field depends on the frequency only, so it is computed ONCE per frequency (one warp
// each) into field[f][0..n) = mu, field[f][n..2n) = mu' and the rays of the fan read it (snell_kernel, p.field).
template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(128) snell_field_kernel(const SnellParams p, int n_freq) {
  const int ins = (p.alt[0] > 0.0) ? 1 : 0;                             // lib:1169
  const int n = p.n_alt + ins;
  const int f = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (f >= n_freq) return;
  double* dst = p.field + (size_t)f * 2 * (size_t)(p.n_alt + 1);
  field_levels<MODE, LITERAL>(p, p.f0_hz[f], n, ins, threadIdx.x & 31, dst, dst + (p.n_alt + 1));
}

template <int MODE, bool SPH, bool LITERAL>
__device__ void trace_one_ray(const SnellParams& p, int64_t ray, const double* s_alt, int n, int ins, WarpRay w) {
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  // fan entry: ray = frequency index * rays_per_freq + elevation index
  const int64_t fi = p.rays_per_freq ? ray / p.rays_per_freq : ray;
  const double f0 = p.f0_hz[fi];
  const double elev = p.elev_deg[p.rays_per_freq ? ray - fi * p.rays_per_freq : ray];
  double* out = p.scalars + ray * 5;
  auto fail = [&]() {                                        // the reference returns NaN for every key
    if (lane < 5) out[lane] = CUDART_NAN;
    if (lane == 0 && p.n_path) p.n_path[ray] = 0;
    if (p.x_out) {
      for (int k = lane; k < p.path_stride; k += 32) {
        p.x_out[ray * (int64_t)p.path_stride + k] = CUDART_NAN;
        p.z_out[ray * (int64_t)p.path_stride + k] = CUDART_NAN;
      }
    }
  };

  // ---- A: field ----
  if (p.rays_per_freq) {
    const double* src = p.field + (size_t)fi * 2 * (size_t)(p.n_alt + 1);
    for (int k = lane; k < n; k += 32) {
      w.xu[k] = src[k];
      w.mup[k] = src[(p.n_alt + 1) + k];
    }
  } else {
    field_levels<MODE, LITERAL>(p, f0, n, ins, lane, w.xu, w.mup);
  }
  __syncwarp();
  const double mu0 = w.xu[0];
  const double s0 = sin(__dmul_rn(__dsub_rn(90.0, elev), kDeg2Rad));  // lib:1188-1189
  if (!isfinite(mu0) || (!SPH && !isfinite(s0))) { fail(); return; }
  const double r_e = p.r_e;
  const double pinv = SPH ? __dmul_rn(__dmul_rn(mu0, __dadd_rn(r_e, s_alt[0])), s0) : __dmul_rn(mu0, s0);

  // ---- B: compaction of the levels with finite mu ----
  int nv = 0;
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    const double m = (k < n) ? w.xu[k] : CUDART_NAN;
    const bool ok = (m == m);
    const unsigned bal = __ballot_sync(full, ok);
    if (ok) {
      const int pos = nv + __popc(bal & ((1u << lane) - 1u));
      w.muv[pos] = m;
      w.vidx[pos] = k;
    }
    nv += __popc(bal);
  }
  __syncwarp();
  if (nv < 2) { fail(); return; }

  // ---- C: first crossing and turning altitude ----
  auto q_of = [&](int i) -> double {
    return SPH ? __dmul_rn(w.muv[i], __dadd_rn(r_e, s_alt[w.vidx[i]])) : w.muv[i];
  };
  int i0 = -1;
  for (int b0 = 0; b0 < nv - 1 && i0 < 0; b0 += 32) {
    const int i = b0 + lane;
    const bool hit = (i < nv - 1) && (q_of(i) >= pinv) && (q_of(i + 1) <= pinv);
    const unsigned bal = __ballot_sync(full, hit);
    if (bal) i0 = b0 + __ffs(bal) - 1;
  }
  if (i0 < 0) { fail(); return; }
  const double qa0 = q_of(i0), qb0 = q_of(i0 + 1);
  const double z0 = s_alt[w.vidx[i0]], z1 = s_alt[w.vidx[i0 + 1]];
  double t = (qa0 == qb0) ? 0.0 : __ddiv_rn(__dsub_rn(qa0, pinv), __dsub_rn(qa0, qb0));
  if (SPH) t = fmin(fmax(t, 0.0), 1.0);                                 // lib:1622
  const double z_turn = (!SPH && qa0 == qb0) ? z0 : __dadd_rn(z0, __dmul_rn(t, __dsub_rn(z1, z0)));
  int n_lev;                                                           // profile levels on the up-leg, apex excluded
  if (SPH) {
    n_lev = i0 + 1;                                                    // lib:1626
  } else {
    int lo = 0, hi = nv;                                               // np.searchsorted(zv, z_turn) (lib:1228)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_alt[w.vidx[mid]] < z_turn) lo = mid + 1; else hi = mid;
    }
    n_lev = lo;
  }
  const int n_up = n_lev + 1;
  const double mu_apex = SPH ? __ddiv_rn(pinv, __dadd_rn(r_e, z_turn)) : pinv;
  auto z_up = [&](int k) -> double { return (k < n_lev) ? s_alt[w.vidx[k]] : z_turn; };
  auto mu_up = [&](int k) -> double { return (k < n_lev) ? w.muv[k] : mu_apex; };

  // ---- D: horizontal coordinate of the up-leg ----
  __syncwarp();
  for (int k = lane; k < n_up - 1; k += 32) {
    const double za = z_up(k), zb = z_up(k + 1);
    const double dz = __dsub_rn(zb, za);
    const double ma = mu_up(k), mb = mu_up(k + 1);
    double term;
    if (!SPH) {
      double mid = __dmul_rn(0.5, __dadd_rn(ma, mb));
      if (k == n_up - 2) mid = fmax(mid, __dadd_rn(pinv, 1e-8));        // lib:1237
      double arg = __dsub_rn(__dmul_rn(mid, mid), __dmul_rn(pinv, pinv));
      if (arg < 1e-10) arg = 1e-10;                                     // lib:1056-1060
      term = __dmul_rn(dz, __ddiv_rn(pinv, __dsqrt_rn(arg)));
    } else if (dz <= 0.0) {
      term = 0.0;                                                       // skipped interval (lib:1646-1647)
    } else {
      const double ra = __dadd_rn(r_e, za), rb = __dadd_rn(r_e, zb);
      const double qa = __dmul_rn(ma, ra), qb = __dmul_rn(mb, rb);
      int nsub = max(1, (int)ceil(__ddiv_rn(fabs(dz), p.dz_target)));   // lib:1650
      const double sharp = __ddiv_rn(1.0, fmin(fmax(__dsub_rn(qa, pinv), 1e-12), fmax(__dsub_rn(qb, pinv), 1e-12)));
      nsub = (int)fmin((double)p.max_substeps,
                       __dmul_rn((double)nsub, __dadd_rn(1.0, __dmul_rn(p.apex_boost, sharp))));   // lib:1656
      const double dn = (double)nsub;
      const double h = __ddiv_rn(dz, dn);
      const double inv_n = __ddiv_rn(1.0, dn);
      const double dmu = __dsub_rn(mb, ma);
      const double p2 = __dmul_rn(pinv, pinv);
      const double ph = __dmul_rn(pinv, h);
      double acc = 0.0;
      // midpoint rule, sequential as lib:1660-1673.  The cancelling difference (mu r)^2 - p^2 keeps the reference's
      // operations; the midpoint parameter (j + 1/2) / N and the quotient p / (r sqrt(.)) go through reciprocal seeds
      // (<= 2 ulp from the reference's three IEEE divisions and square root per sub-step, which made the kernel
      // division-bound: up to 400 sub-steps per level next to the apex).
      for (int j = 0; j < nsub; ++j) {
        const double tm = __dmul_rn((double)j + 0.5, inv_n);
        const double rm = __dadd_rn(r_e, __dadd_rn(za, __dmul_rn(tm, dz)));
        double qm = __dmul_rn(__dadd_rn(ma, __dmul_rn(dmu, tm)), rm);
        if (qm <= pinv) qm = __dadd_rn(pinv, 1e-8);
        const double den = fmax(__dsub_rn(__dmul_rn(qm, qm), p2), 1e-16);
        acc = __dadd_rn(acc, __dmul_rn(ph, __dmul_rn(rsqrt_fast(den), rcp_fast(rm))));
      }
      term = acc;
    }
    w.xu[k + 1] = term;
  }
  __syncwarp();
  if (lane == 0) {                                                      // np.cumsum / phi_up recurrence, sequential
    w.xu[0] = 0.0;
    for (int k = 0; k < n_up - 1; ++k) {
      const bool skipped = SPH && !(__dsub_rn(z_up(k + 1), z_up(k)) > 0.0);
      w.xu[k + 1] = skipped ? 0.0 : __dadd_rn(w.xu[k], w.xu[k + 1]);
    }
  }
  __syncwarp();

  // ---- E: mirrored path, lengths, group delay ----
  const int n_full = 2 * n_up - 1;
  const double h_turn = w.xu[n_up - 1];
  auto hc_full = [&](int k) -> double {                                 // x_up / phi_up mirrored about the apex
    return (k < n_up) ? w.xu[k] : __dsub_rn(__dmul_rn(2.0, h_turn), w.xu[n_full - 1 - k]);
  };
  auto z_full = [&](int k) -> double { return z_up(k < n_up ? k : n_full - 1 - k); };
  auto x_full = [&](int k) -> double { return SPH ? __dmul_rn(r_e, hc_full(k)) : hc_full(k); };
  // mu' on the path: node values at the profile levels, numpy interpolation at the apex (lib:1253 / lib:1691)
  const double mup_apex = np_interp_at(z_turn, np_bracket(z_turn, s_alt, n), s_alt, w.mup, n);
  auto mup_full = [&](int k) -> double {
    const int u = (k < n_up) ? k : n_full - 1 - k;
    return (u < n_lev) ? w.mup[w.vidx[u]] : mup_apex;
  };
  auto seg_len = [&](int s) -> double {
    const double dz = __dsub_rn(z_full(s + 1), z_full(s));
    if (SPH) {
      const double rmid = __dadd_rn(r_e, __dmul_rn(0.5, __dadd_rn(z_full(s), z_full(s + 1))));
      return hypot(__dmul_rn(rmid, __dsub_rn(hc_full(s + 1), hc_full(s))), dz);
    }
    return hypot(__dsub_rn(x_full(s + 1), x_full(s)), dz);
  };
  double path = 0.0, delay = 0.0;
  for (int s = lane; s < n_full - 1; s += 32) {
    const double ds = seg_len(s);
    if (s < n_up - 1) w.muv[s] = ds;                                    // up-leg lengths for the midpoint search
    if (ds == ds) path += ds;                                           // nansum
    const double term = __dmul_rn(__ddiv_rn(__dmul_rn(0.5, __dadd_rn(mup_full(s + 1), mup_full(s))), kCkmS), ds);
    if (term == term) delay += term;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    path += __shfl_xor_sync(full, path, o);
    delay += __shfl_xor_sync(full, delay, o);
  }
  __syncwarp();
  if (p.x_out) {
    double* xo = p.x_out + ray * (int64_t)p.path_stride;
    double* zo = p.z_out + ray * (int64_t)p.path_stride;
    for (int k = lane; k < p.path_stride; k += 32) {
      xo[k] = (k < n_full) ? x_full(k) : CUDART_NAN;
      zo[k] = (k < n_full) ? z_full(k) : CUDART_NAN;
    }
  }
  if (lane == 0) {
    double xm = CUDART_NAN, zm = CUDART_NAN;
    if (path > 0.0) {                                                   // lib:1258-1264
      const double half = __dmul_rn(0.5, path);
      double s_cum = 0.0;
      int mid = n_full - 1;                                             // searchsorted returns len(s_cum) when never reached
      for (int s = 0; s < n_full - 1; ++s) {
        s_cum = __dadd_rn(s_cum, (s < n_up - 1) ? w.muv[s] : seg_len(s));
        if (s_cum >= half) { mid = s; break; }
      }
      xm = x_full(mid);
      zm = z_full(mid);
    }
    out[0] = path;
    out[1] = delay;
    out[2] = xm;
    out[3] = zm;
    out[4] = (fabs(z_full(n_full - 1)) <= 1e-3) ? x_full(n_full - 1) : CUDART_NAN;    // lib:1266-1268
    if (p.n_path) p.n_path[ray] = n_full;
  }
  __syncwarp();
}

template <int MODE, bool SPH, bool LITERAL>
__global__ void __launch_bounds__(128) snell_kernel(const SnellParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ins = (p.alt[0] > 0.0) ? 1 : 0;                             // lib:1169
  const int n = p.n_alt + ins;
  double* s_alt = reinterpret_cast<double*>(smem_raw);
  for (int k = threadIdx.x; k < n; k += blockDim.x) s_alt[k] = (ins && k == 0) ? 0.0 : p.alt[k - ins];
  __syncthreads();
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5;
  const size_t per_warp = (size_t)(3 * n + 1) * sizeof(double) + (size_t)((n + 1) & ~1) * sizeof(int);
  unsigned char* base = smem_raw + (size_t)((n + 1) & ~1) * sizeof(double) + (size_t)wid * per_warp;
  WarpRay w;
  w.muv = reinterpret_cast<double*>(base);
  w.mup = w.muv + n;
  w.xu = w.mup + n;
  w.vidx = reinterpret_cast<int*>(w.xu + n + 1);
  for (int64_t ray = (int64_t)blockIdx.x * warps + wid; ray < p.n_rays; ray += (int64_t)gridDim.x * warps) {
    trace_one_ray<MODE, SPH, LITERAL>(p, ray, s_alt, n, ins, w);
    __syncwarp();
  }
}

template <int MODE, bool SPH, bool LITERAL>
cudaError_t launch_snell_t(const SnellParams& p, int max_smem_optin, cudaStream_t stream) {
  const int n = p.n_alt + 1;
  const size_t shared_alt = (size_t)((n + 1) & ~1) * sizeof(double);
  const size_t per_warp = (size_t)(3 * n + 1) * sizeof(double) + (size_t)((n + 1) & ~1) * sizeof(int);
  int warps = 4;
  while (warps > 1 && shared_alt + warps * per_warp > (size_t)max_smem_optin) --warps;
  const size_t smem = shared_alt + warps * per_warp;
  if (smem > (size_t)max_smem_optin) return cudaErrorInvalidValue;
  auto kern = snell_kernel<MODE, SPH, LITERAL>;
  cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t ctas = (p.n_rays + warps - 1) / warps;
  if (ctas > 148 * 16) ctas = 148 * 16;
  if (p.rays_per_freq) {
    const int n_freq = (int)(p.n_rays / p.rays_per_freq);
    snell_field_kernel<MODE, LITERAL><<<(unsigned)((n_freq + 3) / 4), 128, 0, stream>>>(p, n_freq);
  }
  kern<<<(unsigned)ctas, warps * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

size_t snell_smem_bytes(int n_alt) {
  const int n = n_alt + 1;
  return (size_t)((n + 1) & ~1) * sizeof(double) + (size_t)(3 * n + 1) * sizeof(double) +
         (size_t)((n + 1) & ~1) * sizeof(int);
}

cudaError_t launch_snell(const SnellParams& p, int max_smem_optin, cudaStream_t stream) {
  if (p.n_rays <= 0) return cudaSuccess;
#define PRHF_SNELL(M, S, L) return launch_snell_t<M, S, L>(p, max_smem_optin, stream)
  if (p.mode == 0) {
    if (p.spherical) { if (p.literal) PRHF_SNELL(0, true, true); else PRHF_SNELL(0, true, false); }
    else { if (p.literal) PRHF_SNELL(0, false, true); else PRHF_SNELL(0, false, false); }
  } else {
    if (p.spherical) { if (p.literal) PRHF_SNELL(1, true, true); else PRHF_SNELL(1, true, false); }
    else { if (p.literal) PRHF_SNELL(1, false, true); else PRHF_SNELL(1, false, false); }
  }
#undef PRHF_SNELL
}

}  // namespace prhf

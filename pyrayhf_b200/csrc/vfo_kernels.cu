// Fused vertical-forward-operator kernels for B200 (sm_100a).
//
// One CTA evaluates one tile = (profile, sounding frequency, segment of the stretched grid).
// Everything the reference materialises as [n_freq x n_points] arrays (lib:410-438 new_alt, dist,
// den/bmag/bpsi on the grid; lib:500-503 X, Y; ~40 temporaries in lib:209-254) lives in registers;
// the profile lives in shared memory; HBM sees the inputs once and one double per virtual height.
//
// Stages inside the CTA (reference lines in PyRayHF/library.py):
//   1. stage den/alt, argmax(den) -> truncation below the peak          lib:371-375
//   2. node tables (slopes for np.interp, sin/cos of the field angle)    lib:424-426, lib:210-211
//   3. critical curve X or X+Y at the nodes, running max, validity,
//      reflection height by np.interp(1.0, ...), back-off 1e-6 km        lib:380-407
//   4. stretched grid h_i = m_i (h_c - alt0) + alt0, dh_i = h_{i+1}-h_i  lib:413-416
//   5. interpolate, X, Y, Appleton-Hartree mu', left-Riemann sum         lib:424-426, 500-506, 209-254, 288
//   6. block reduction, ==0 -> NaN, + min(alt)                           lib:288-292
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "vfo_device.cuh"
#include "vfo_kernels.h"

namespace prhf {

// ------------------------------------------------------------------------------------------
// stretched-grid multiplier table (lib:314-320): m_i = 1 - (exp(10 (1-u_i)) - 1)/(exp(10) - 1)
// ------------------------------------------------------------------------------------------
__global__ void grid_multiplier_kernel(int n, double step, double* __restrict__ m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u = __dmul_rn((double)i, step);        // np.linspace: arange(n) * step ...
  if (i == n - 1 && n > 1) u = 1.0;             // ... with the endpoint forced
  const double fl = __dsub_rn(1.0, u);
  const double den = __dsub_rn(exp(kSharp), 1.0);
  const double factor = __ddiv_rn(__dsub_rn(exp(__dmul_rn(kSharp, fl)), 1.0), den);
  m[i] = __dsub_rn(1.0, factor);
}

// ------------------------------------------------------------------------------------------
// block-wide helpers (kThreads = 256 = 8 warps)
// ------------------------------------------------------------------------------------------
struct BlockScratch {
  double d[kThreads / 32];
  int i[kThreads / 32];
  double bcast_d[4];
  int bcast_i[4];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the block in a fixed order; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, BlockScratch& sc) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sc.d[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) r += sc.d[k];
  }
  return r;
}
// max / min over the block, broadcast to every thread (fmax/fmin ignore NaN)
__device__ __forceinline__ double block_max(double v, BlockScratch& sc) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sc.d[wid] = v;
  __syncthreads();
  double r = sc.d[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) r = fmax(r, sc.d[k]);
  return r;
}
__device__ __forceinline__ double block_min(double v, BlockScratch& sc) {
  v = warp_min(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sc.d[wid] = v;
  __syncthreads();
  double r = sc.d[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) r = fmin(r, sc.d[k]);
  return r;
}
__device__ __forceinline__ int block_min_i(int v, BlockScratch& sc) {
  v = warp_min_i(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sc.i[wid] = v;
  __syncthreads();
  int r = sc.i[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) r = min(r, sc.i[k]);
  return r;
}

// np.argmax ordering: NaN wins, then larger value, then first occurrence.
__device__ __forceinline__ bool arg_precedes(double av, int ai, double bv, int bi) {
  const bool an = isnan(av), bn = isnan(bv);
  if (an != bn) return an;
  if (an) return ai < bi;
  if (av != bv) return av > bv;
  return ai < bi;
}
__device__ __forceinline__ int block_argmax(double v, int idx, BlockScratch& sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (arg_precedes(ov, oi, v, idx)) { v = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { sc.d[wid] = v; sc.i[wid] = idx; }
  __syncthreads();
  double bv = sc.d[0];
  int bi = sc.i[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k)
    if (arg_precedes(sc.d[k], sc.i[k], bv, bi)) { bv = sc.d[k]; bi = sc.i[k]; }
  return bi;
}

// numpy binary_search_with_guess outcome restricted to [lo, hi]: last j in [lo, hi] with xp[j] <= x,
// lo - 1 when x < xp[lo].  (The caller guarantees the true bracket lies in [lo - 1, hi].)
__device__ __forceinline__ int bracket_in(double x, const double* xp, int lo, int hi) {
  int a = lo, b = hi + 1;
  while (a < b) {
    const int mid = a + ((b - a) >> 1);
    if (x >= xp[mid]) a = mid + 1; else b = mid;
  }
  return a - 1;
}

struct NodeTables {
  const double* alt;    // [nt]
  const double* den;    // [nt]
  const double* b;      // [nt]
  const double* psi;    // [nt] degrees
  const double* sden;   // slopes (last entry 0)
  const double* sb;
  const double* spsi;   // degrees / km
  const double* srad;   // radians / km
  const double* sn;     // sin(psi_k)
  const double* cs;     // cos(psi_k)
};

struct RowConst {
  double f_hz;      // lib:491
  double kx;        // cp^2 / f^2     (fast path: X = den * kx)
  double ky;        // g_p / f        (fast path: Y = b * ky)
  double alt0;      // aalt[0]
  double span;      // h_c - aalt[0]  (lib:413)
  int nt;           // truncated length (= argmax(den))
  int jlo, jhi;     // node window covered by this tile
  bool degenerate;  // h_c < alt0: every grid point clamps to node 0
};

// One tile of grid points [i0, i1); returns this thread's partial nansum of mu' * dh.
// GENERAL: numpy-literal interpolation + libdevice sincos (needed for non-finite node values or large
//          per-segment angle steps).  LITERAL: additionally the reference-order Appleton-Hartree.
template <int MODE, bool GENERAL, bool LITERAL, bool ISO>
__device__ __forceinline__ double tile_sum(const NodeTables& T, const RowConst& rc, const double* __restrict__ m,
                                           int i0, int i1, int n_points) {
  double acc = 0.0;
  for (int i = i0 + (int)threadIdx.x; i < i1; i += kThreads) {
    const double mi = __ldg(m + i);
    const double h = __dadd_rn(__dmul_rn(mi, rc.span), rc.alt0);                 // lib:413
    double dh;
    if (i + 1 < n_points) {
      const double hn = __dadd_rn(__dmul_rn(__ldg(m + i + 1), rc.span), rc.alt0);
      dh = __dsub_rn(hn, h);                                                     // lib:415
    } else {
      dh = kBackoff;                                                             // lib:416
    }
    int j = bracket_in(h, T.alt, rc.jlo, rc.jhi);
    double mup;
    if (GENERAL || LITERAL) {
      const double den = interp_numpy(h, j, rc.nt, T.alt, T.den, T.sden);        // lib:424
      const double b = interp_numpy(h, j, rc.nt, T.alt, T.b, T.sb);              // lib:425
      const double X = x_literal(den, rc.f_hz);                                  // lib:500
      if (ISO) {
        mup = iso_mup(X, nullptr);
      } else {
        const double psi = interp_numpy(h, j, rc.nt, T.alt, T.psi, T.spsi);      // lib:426
        const double Y = y_literal(b, rc.f_hz);                                  // lib:503
        if (LITERAL) {
          mup = ah_literal<MODE>(X, Y, psi, nullptr);
        } else {
          double sn, cs;
          sincos(__dmul_rn(psi, kDeg2Rad), &sn, &cs);
          mup = ah_fast<MODE>(X, Y, sn, cs, nullptr);
        }
      }
    } else {
      j = max(j, 0);
      const double t = rc.degenerate ? 0.0 : (h - T.alt[j]);
      const double X = fma(T.sden[j], t, T.den[j]) * rc.kx;
      if (ISO) {
        mup = iso_mup(X, nullptr);
      } else {
        const double Y = fma(T.sb[j], t, T.b[j]) * rc.ky;
        double sn, cs;
        rotate_sincos(T.sn[j], T.cs[j], T.srad[j] * t, &sn, &cs);
        mup = ah_fast<MODE>(X, Y, sn, cs, nullptr);
      }
    }
    const double term = mup * dh;                                                // lib:288
    acc += (term == term) ? term : 0.0;                                          // nansum
  }
  return acc;
}

template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kThreads, 2) vfo_tile_kernel(const VfoParams p) {
  extern __shared__ __align__(16) double smem[];
  __shared__ BlockScratch sc;
  const int A = p.n_alt;
  double* s_alt = smem;
  double* s_den = s_alt + A;
  double* s_b = s_den + A;
  double* s_psi = s_b + A;
  double* s_sden = s_psi + A;
  double* s_sb = s_sden + A;
  double* s_spsi = s_sb + A;
  double* s_srad = s_spsi + A;
  double* s_sn = s_srad + A;
  double* s_cs = s_sn + A;
  double* s_crit = s_cs + A;

  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int seg = (int)(tile % p.n_seg);
  const int64_t row = tile / p.n_seg;
  const int r = (int)(row % p.n_freq);
  const int64_t prof = p.profile_offset + row / p.n_freq;
  const int64_t out_idx = prof * p.n_freq + r;

  const double* g_den = p.den + prof * A;
  const double* g_b = p.bmag + prof * A;
  const double* g_psi = p.bpsi + prof * A;
  const double* g_alt = p.alt + prof * p.alt_stride;
  const double f_mhz = p.freq[prof * p.freq_stride + r];

  // ---- 1. stage density + altitude, argmax(den) (lib:371), min(alt) (lib:507) ----
  double best_v = -CUDART_INF;
  int best_i = 0x7fffffff;
  double amin = CUDART_INF;
  for (int k = tid; k < A; k += kThreads) {
    const double d = g_den[k];
    const double a = g_alt[k];
    s_den[k] = d;
    s_alt[k] = a;
    if (arg_precedes(d, k, best_v, best_i)) { best_v = d; best_i = k; }
    amin = fmin(amin, a);
  }
  const int nt = block_argmax(best_v, best_i, sc);      // truncated length = index of the peak
  const double alt_min = block_min(amin, sc);

  if (nt == 0) {                                         // IndexError in the reference (lib:399)
    if (seg == 0 && tid == 0) {
      p.vh[out_idx] = CUDART_NAN;
      if (r == 0 && p.status) p.status[prof] = 2;
    }
    return;
  }

  // ---- 2. node tables over [0, nt) ----
  bool neg = false, nonfinite = false, bigstep = false;
  double bmax = 0.0;
  for (int k = tid; k < nt; k += kThreads) {
    const double d = s_den[k];
    const double b = g_b[k];
    const double ps = g_psi[k];
    s_b[k] = b;
    s_psi[k] = ps;
    neg |= (d < 0.0);
    nonfinite |= !(isfinite(d) && isfinite(b) && isfinite(ps) && isfinite(s_alt[k]));
    bmax = fmax(bmax, fabs(b));
    double sn, cs;
    sincos(ps * kDeg2Rad, &sn, &cs);
    s_sn[k] = sn;
    s_cs[k] = cs;
  }
  if (__syncthreads_or(neg)) {                           // ValueError in the reference (lib:93-94)
    if (seg == 0 && tid == 0) {
      p.vh[out_idx] = CUDART_NAN;
      if (r == 0 && p.status) p.status[prof] = 1;
    }
    return;
  }
  for (int k = tid; k < nt; k += kThreads) {
    double sd = 0.0, sb = 0.0, sp = 0.0;
    if (k + 1 < nt) {                                    // numpy: slopes[k] = (fp[k+1]-fp[k])/(xp[k+1]-xp[k])
      const double dx = __dsub_rn(s_alt[k + 1], s_alt[k]);
      sd = __ddiv_rn(__dsub_rn(s_den[k + 1], s_den[k]), dx);
      sb = __ddiv_rn(__dsub_rn(s_b[k + 1], s_b[k]), dx);
      sp = __ddiv_rn(__dsub_rn(s_psi[k + 1], s_psi[k]), dx);
      nonfinite |= !(dx > 0.0);
      bigstep |= !(fabs(__dsub_rn(s_psi[k + 1], s_psi[k])) * kDeg2Rad <= kMaxRotateStep);
    }
    s_sden[k] = sd;
    s_sb[k] = sb;
    s_spsi[k] = sp;
    s_srad[k] = sp * kDeg2Rad;
  }
  const bool general = __syncthreads_or(nonfinite || bigstep);
  if (seg == 0 && r == 0 && tid == 0 && p.status) p.status[prof] = 0;

  // Unmagnetised switch (lib:201), decided per profile from the node values: isotropic iff
  // g_p * max|B| / min|f| < 1e-12 over the profile's frequencies.  (The reference takes nanmax|Y|
  // over the regridded [F x N] array of one call; the two differ only for |B| ~ 1e-17 T.)
  bool iso = false;
  bmax = block_max(bmax, sc);
  if (bmax < 1e-9) {
    double fmin_abs = CUDART_INF;
    for (int k = tid; k < p.n_freq; k += kThreads) {
      const double f = fabs(__dmul_rn(p.freq[prof * p.freq_stride + k], 1e6));
      if (f > 0.0) fmin_abs = fmin(fmin_abs, f);
    }
    fmin_abs = block_min(fmin_abs, sc);
    iso = (bmax == 0.0) || (__ddiv_rn(__dmul_rn(kGp, bmax), fmin_abs) < kYTol);
  }

  // ---- 3. critical curve at the nodes, validity, reflection height (lib:380-407) ----
  const double f_hz = __dmul_rn(f_mhz, 1e6);
  int first_gt = 0x7fffffff;
  bool has_nan = false, any_ge = false;
  for (int k = tid; k < nt; k += kThreads) {
    double v = x_literal(s_den[k], f_hz);
    if (MODE == 1) v = __dadd_rn(v, y_literal(s_b[k], f_hz));
    s_crit[k] = v;
    has_nan |= isnan(v);
    any_ge |= (v >= 1.0);
    if (v > 1.0) first_gt = min(first_gt, k);
  }
  const int jstar = block_min_i(first_gt, sc);
  const bool dead_nan = __syncthreads_or(has_nan);
  const bool reach = __syncthreads_or(any_ge);
  if (dead_nan || !reach) {                               // valid == False (lib:399) -> NaN (lib:407)
    if (seg == 0 && tid == 0) {
      double res = CUDART_NAN;
      if (nt == 1) {
        // numpy's single-node np.interp has no NaN test: interp(NaN, [x0], [f0]) == f0.  A dead row of a
        // one-node profile therefore still sees finite den/bmag/bpsi, every dh is NaN except the final
        // 1e-6 (lib:416), and the reference returns alt_min + mu'(node 0) * 1e-6.
        const double X = x_literal(s_den[0], f_hz);
        double mup;
        if (iso) mup = iso_mup(X, nullptr);
        else if (LITERAL) mup = ah_literal<MODE>(X, y_literal(s_b[0], f_hz), s_psi[0], nullptr);
        else mup = ah_fast<MODE>(X, y_literal(s_b[0], f_hz), s_sn[0], s_cs[0], nullptr);
        const double term = mup * kBackoff;
        if (term == term && term != 0.0) res = term + alt_min;
      }
      p.vh[out_idx] = res;
    }
    return;
  }
  double hcrit;
  if (jstar == 0 || nt == 1) {
    hcrit = s_alt[0];                                     // 1.0 < fcrit[0]: np.interp clamps left
  } else if (jstar == 0x7fffffff) {
    hcrit = s_alt[nt - 1];                                // running max ends exactly at 1.0
  } else {
    double pm = -CUDART_INF;
    for (int k = tid; k < jstar; k += kThreads) pm = fmax(pm, s_crit[k]);
    const double M = block_max(pm, sc);                   // cummax[jstar-1]
    const int j = jstar - 1;
    if (M == 1.0) {
      hcrit = s_alt[j];
    } else {
      const double slope = __ddiv_rn(__dsub_rn(s_alt[j + 1], s_alt[j]), __dsub_rn(s_crit[jstar], M));
      hcrit = __dadd_rn(__dmul_rn(slope, __dsub_rn(1.0, M)), s_alt[j]);
    }
  }
  const double hc = __dsub_rn(hcrit, kBackoff);           // lib:407

  RowConst rc;
  rc.f_hz = f_hz;
  rc.kx = (kCp * kCp) / (f_hz * f_hz);
  rc.ky = kGp / f_hz;
  rc.alt0 = s_alt[0];
  rc.span = __dsub_rn(hc, s_alt[0]);
  rc.nt = nt;
  rc.degenerate = (hc < s_alt[0]) || (nt == 1);

  // ---- 4. node window of this tile ----
  const int i0 = seg * p.seg_len;
  const int i1 = min(p.n_points, i0 + p.seg_len);
  __syncthreads();
  if (tid == 0 || tid == 32) {
    const int i = (tid == 0) ? i0 : (i1 - 1);
    const double h = __dadd_rn(__dmul_rn(__ldg(p.mult + i), rc.span), rc.alt0);
    sc.bcast_i[tid >> 5] = bracket_in(h, s_alt, 0, nt - 1);
  }
  __syncthreads();
  {
    const int ja = sc.bcast_i[0], jb = sc.bcast_i[1];
    rc.jlo = max(min(ja, jb), 0);
    rc.jhi = max(max(ja, jb), 0);
  }

  NodeTables T{s_alt, s_den, s_b, s_psi, s_sden, s_sb, s_spsi, s_srad, s_sn, s_cs};

  // ---- 5. grid points of the tile ----
  double acc;
  if (LITERAL) {
    acc = iso ? tile_sum<MODE, true, true, true>(T, rc, p.mult, i0, i1, p.n_points)
              : tile_sum<MODE, true, true, false>(T, rc, p.mult, i0, i1, p.n_points);
  } else if (iso) {
    acc = tile_sum<MODE, true, false, true>(T, rc, p.mult, i0, i1, p.n_points);
  } else if (general) {
    acc = tile_sum<MODE, true, false, false>(T, rc, p.mult, i0, i1, p.n_points);
  } else {
    acc = tile_sum<MODE, false, false, false>(T, rc, p.mult, i0, i1, p.n_points);
  }

  // ---- 6. reduce, finish (lib:288-292) ----
  const double s_tile = block_sum(acc, sc);
  if (tid != 0) return;
  double total = s_tile;
  if (p.n_seg > 1) {
    const int64_t local_row = row;                        // row index inside this launch
    double* part = p.partial + local_row * p.n_seg;
    __stcg(part + seg, s_tile);
    __threadfence();
    const unsigned prev = atomicAdd(p.counter + local_row, 1u);
    if (prev != (unsigned)(p.n_seg - 1)) return;
    __threadfence();
    total = 0.0;
    for (int s = 0; s < p.n_seg; ++s) total += __ldcg(part + s);   // fixed order: deterministic
    p.counter[local_row] = 0u;                            // self-reset for the next launch
  }
  if (total == 0.0) total = CUDART_NAN;                   // lib:290
  p.vh[out_idx] = total + alt_min;                        // lib:292
}

// ------------------------------------------------------------------------------------------
// elementwise mu / mu' (lib:161-256) for callers outside the fused path
// ------------------------------------------------------------------------------------------
template <int MODE, bool LITERAL, bool ISO>
__global__ void mu_mup_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                              const double* __restrict__ psi, int64_t n, double* __restrict__ mu_out,
                              double* __restrict__ mup_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double mu, mup;
    if (ISO) {
      mup = iso_mup(X[i], &mu);
    } else if (LITERAL) {
      mup = ah_literal<MODE>(X[i], Y[i], psi[i], &mu);
    } else {
      double sn, cs;
      sincos(psi[i] * kDeg2Rad, &sn, &cs);
      mup = ah_fast<MODE>(X[i], Y[i], sn, cs, &mu);
    }
    if (mu_out) mu_out[i] = mu;
    if (mup_out) mup_out[i] = mup;
  }
}

// ------------------------------------------------------------------------------------------
// FP64 FMA throughput probe (roofline denominator)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;   // never true; keeps the chain alive
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
size_t vfo_smem_bytes(int n_alt) { return sizeof(double) * 11 * (size_t)n_alt; }

template <int MODE, bool LITERAL>
static cudaError_t launch_tiles(const VfoParams& p, int64_t n_tiles, cudaStream_t stream) {
  const size_t smem = vfo_smem_bytes(p.n_alt);
  auto kern = vfo_tile_kernel<MODE, LITERAL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)n_tiles, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_vfo_tiles(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream) {
  if (mode == 0) return literal ? launch_tiles<0, true>(p, n_tiles, stream) : launch_tiles<0, false>(p, n_tiles, stream);
  return literal ? launch_tiles<1, true>(p, n_tiles, stream) : launch_tiles<1, false>(p, n_tiles, stream);
}

cudaError_t launch_grid_multiplier(int n, double* m, cudaStream_t stream) {
  const double step = (n > 1) ? 1.0 / (double)(n - 1) : 0.0;
  grid_multiplier_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, step, m);
  return cudaGetLastError();
}

cudaError_t launch_mu_mup(const double* X, const double* Y, const double* psi, int64_t n, int mode, bool iso,
                          bool literal, double* mu, double* mup, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
#define PRHF_LAUNCH(M, L, I) mu_mup_kernel<M, L, I><<<(unsigned)blocks, threads, 0, stream>>>(X, Y, psi, n, mu, mup)
  if (iso) PRHF_LAUNCH(0, false, true);
  else if (mode == 0 && literal) PRHF_LAUNCH(0, true, false);
  else if (mode == 0) PRHF_LAUNCH(0, false, false);
  else if (literal) PRHF_LAUNCH(1, true, false);
  else PRHF_LAUNCH(1, false, false);
#undef PRHF_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_dfma_probe(double* out, int blocks, int iters, cudaStream_t stream) {
  dfma_probe_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0);
  return cudaGetLastError();
}

}  // namespace prhf

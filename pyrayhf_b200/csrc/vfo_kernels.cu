// Fused vertical-forward-operator kernels for B200 (sm_100a).
//
// One launch (single profile) or two (batches) per call; no [n_freq x n_points] array ever reaches HBM:
//
//   vfo_rows_kernel  (K1)  row setup, one warp (small batches) or one thread (large batches) per sounding frequency:
//       peak truncation, error status, unmagnetised switch, critical curve X or X+Y at the profile
//       nodes, running max, validity, reflection height, back-off.          lib:371-407
//       Output: 8 bytes per (profile, frequency) row (h_c - alt0, NaN = row finished) and a 64-byte
//       record per profile.  Rows that do not reflect get their NaN here and cost nothing later; on large
//       batches rows clamped to the first level are finished here too and the rest is queued.
//
//   grid points (K2)       one tile = (profile, frequency, segment of the stretched grid):
//       stages the profile nodes the tile touches (np.interp slopes, sin/cos of the field angle) in
//       shared memory, then per grid point: h_i, dh_i (lib:413-416), linear interpolation of
//       den/bmag/bpsi (lib:424-426), X, Y (lib:500-503), Appleton-Hartree mu' (lib:209-254), the
//       left-Riemann nansum (lib:288), block reduction, ==0 -> NaN, + min(alt) (lib:290-292).
//       Everything the reference materialises per grid point lives in registers.
//         vfo_queue_kernel    large batches: 128-thread CTAs, eight per SM, whole-row tiles drawn from K1's queue by ticket
//         vfo_tile_kernel     one CTA per row / planned segments of small batches / rows the queue kernel deferred
//         vfo_rowwarp_kernel  n_points <= 4096: one warp per row, nodes staged once per profile
//         vfo_solo_kernel     one profile: row setup and tile in ONE launch
//         vfo_tile_global_kernel  profiles with more levels than shared memory holds
//       Uniform altitude grids run the grid loop in "E-space" (tile_sum_fast_e): the stretched grid is a geometric
//       sequence, so the loop computes it with one multiplication per point instead of reading a table.
//
// PyRayHF/library.py is abbreviated "lib" throughout.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "vfo_device.cuh"
#include "vfo_kernels.h"

namespace prhf {

// Optional per-CTA phase trace (developer builds only: make TRACE=1).  Thread 0 of every tile-kernel CTA
// stores {smid, globaltimer at entry, clock64 at 6 phase boundaries} into p.trace.
#ifdef PRHF_TRACE
#define PRHF_TRACE_MARK(slot)                                                              \
  do {                                                                                     \
    if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 8 + (slot)] = clock64(); \
  } while (0)
__device__ __forceinline__ long long trace_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int trace_smid() {
  int s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}
#define PRHF_TRACE_X(slot)                                                                                   \
  do {                                                                                                     \
    if (p.trace_k1 && threadIdx.x == 0) p.trace_k1[16384 + (size_t)blockIdx.x * 16 + (slot)] = clock64(); \
  } while (0)
#define PRHF_TRACE_K1(slot)                                                                      \
  do {                                                                                           \
    if (p.trace_k1 && threadIdx.x == 0) p.trace_k1[(size_t)blockIdx.x * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define PRHF_TRACE_MARK(slot) do { } while (0)
#define PRHF_TRACE_K1(slot) do { } while (0)
#define PRHF_TRACE_X(slot) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// stretched-grid multiplier table (lib:314-320): m_i = 1 - (exp(10 (1-u_i)) - 1)/(exp(10) - 1)
// The table has kMultPad extra entries (copies of the last entry) so that the main loop can read m[i+1],
// m[i+2] unconditionally and the row's last point gets the in-loop weight h(pad) - h(last) = 0.
// ------------------------------------------------------------------------------------------
// dm_i = m_{i+1} - m_i: the weight of grid point i in units of the row's span (lib:415: dh_i = h_{i+1} - h_i with
// h = m (h_c - alt0) + alt0, lib:413); 0 from the row's last point on (that point weighs 1e-6 km, lib:416, added by
// its owner after the loop).
__global__ void grid_dmult_kernel(int n, int n_padded, const double* __restrict__ m, double* __restrict__ dm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  dm[i] = (i < n - 1) ? __dsub_rn(m[i + 1], m[i]) : 0.0;
}

// e (optional): E_i = exp(10 (1 - u_i)), the exponential inside m_i; the E-space grid loop (tile_sum_fast_e) seeds its
// recurrence from it.  Entries past n_points are 1 (= E at the row's last point).
__global__ void grid_multiplier_kernel(int n, int n_padded, double step, double* __restrict__ m,
                                       double* __restrict__ e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded) return;
  if (i >= n) {                                  // copies of the last entry m[n-1]
    m[i] = (n > 1) ? 1.0 : 0.0;
    if (e) e[i] = (n > 1) ? 1.0 : exp(kSharp);
    return;
  }
  double u = __dmul_rn((double)i, step);        // np.linspace: arange(n) * step ...
  if (i == n - 1 && n > 1) u = 1.0;             // ... with the endpoint forced
  const double fl = __dsub_rn(1.0, u);
  const double den = __dsub_rn(exp(kSharp), 1.0);
  const double ex = exp(__dmul_rn(kSharp, fl));
  const double factor = __ddiv_rn(__dsub_rn(ex, 1.0), den);
  m[i] = __dsub_rn(1.0, factor);
  if (e) e[i] = ex;
}

// ------------------------------------------------------------------------------------------
// warp / block helpers (any block size that is a multiple of 32, at most kMaxWarps warps)
// ------------------------------------------------------------------------------------------
struct BlockScratch {
  double d[kMaxWarps];
  int i[kMaxWarps];
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
    for (int k = 0; k < kMaxWarps; ++k) r += (k < (int)(blockDim.x >> 5)) ? sc.d[k] : 0.0;
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
  for (int k = 1; k < kMaxWarps; ++k) if (k < (int)(blockDim.x >> 5)) r = fmax(r, sc.d[k]);
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
  for (int k = 1; k < kMaxWarps; ++k) if (k < (int)(blockDim.x >> 5)) r = fmin(r, sc.d[k]);
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
  for (int k = 1; k < kMaxWarps; ++k)
    if (k < (int)(blockDim.x >> 5) && arg_precedes(sc.d[k], sc.i[k], bv, bi)) { bv = sc.d[k]; bi = sc.i[k]; }
  return bi;
}

// ---- K1 profile reductions: one shared-memory round each ----
// Total order key for np.argmax: NaN above everything, then the value; ties broken by the lower index.
__device__ __forceinline__ long long argmax_key(double v) {
  long long b = __double_as_longlong(v);
  b = (b < 0) ? (b ^ 0x7fffffffffffffffLL) : b;           // sign-magnitude -> two's-complement order
  return isnan(v) ? 0x7fffffffffffffffLL : b;
}
struct ProfileReduce1 { int nt; double alt_min; };
__device__ __forceinline__ ProfileReduce1 block_argmax_min(double v, int idx, double amin, long long* s_key,
                                                           int* s_idx, double* s_min) {
  long long key = argmax_key(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    const double om = __shfl_xor_sync(0xffffffffu, amin, o);
    if (ok > key || (ok == key && oi < idx)) { key = ok; idx = oi; }
    amin = fmin(amin, om);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_key[wid] = key; s_idx[wid] = idx; s_min[wid] = amin; }
  __syncthreads();
  ProfileReduce1 r;
  long long bk = s_key[0];
  r.nt = s_idx[0];
  r.alt_min = s_min[0];
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) {
    if (s_key[k] > bk || (s_key[k] == bk && s_idx[k] < r.nt)) { bk = s_key[k]; r.nt = s_idx[k]; }
    r.alt_min = fmin(r.alt_min, s_min[k]);
  }
  return r;
}
struct ProfileReduce2 { double bmax, step_max; int flags; };
__device__ __forceinline__ ProfileReduce2 block_max2_or(double a, double b, int flags, double* s_a, double* s_b,
                                                        int* s_f) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    flags |= __shfl_xor_sync(0xffffffffu, flags, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_a[wid] = a; s_b[wid] = b; s_f[wid] = flags; }
  __syncthreads();
  ProfileReduce2 r{s_a[0], s_b[0], s_f[0]};
#pragma unroll
  for (int k = 1; k < kThreads / 32; ++k) {
    r.bmax = fmax(r.bmax, s_a[k]);
    r.step_max = fmax(r.step_max, s_b[k]);
    r.flags |= s_f[k];
  }
  return r;
}

// numpy binary_search_with_guess outcome restricted to [lo, hi]: last j in [lo, hi] with xp[j] <= x,
// lo - 1 when x < xp[lo].  STRIDE is in doubles (the node table interleaves 8 doubles per node).
template <int STRIDE>
__device__ __forceinline__ int bracket_in(double x, const double* xp, int lo, int hi) {
  int a = lo, b = hi + 1;
  while (a < b) {
    const int mid = a + ((b - a) >> 1);
    if (x >= xp[(ptrdiff_t)mid * STRIDE]) a = mid + 1; else b = mid;
  }
  return a - 1;
}

// ==========================================================================================
// K1: per-row setup
// ==========================================================================================
// Result of the critical-curve scan of one row.
struct RowScan {
  int jstar;        // first level whose literal value exceeds 1 (0x7fffffff: none)
  double v_jstar;   // its literal value
  double M;         // literal running max just below it
  bool any_eq1;     // some literal value equals 1 exactly (matters only when jstar is none)
  bool has_nan;
};

// Warp-per-row scan:
// screen == true: iterate over screen candidates until a literal value exceeds 1, then evaluate literally
// every level whose screen value is within tolerance of the running max; screen == false (non-finite
// inputs): literal values at every level.
__device__ __forceinline__ RowScan row_scan_general(bool screen, int mode, int nt, int lane, double f_hz, double kx,
                                                 double ky, const double* s_den, const double* s_b, double* crit) {
  int jstar = 0x7fffffff;
  bool any_eq1 = false, has_nan = false;
  double v_jstar = 0.0, M = -CUDART_INF;
  if (screen) {
      for (int k = lane; k < nt; k += 32) crit[k] = fma(s_b[k], ky, s_den[k] * kx);
      __syncwarp();
      int start = 0;
      for (;;) {
        int cand = 0x7fffffff;
        for (int k = lane; k < nt; k += 32) {
          if (k < start) continue;
          const double tol = kScreenTol * (fabs(s_den[k] * kx) + fabs(s_b[k] * ky));
          if (crit[k] >= 1.0 - tol) { cand = k; break; }
        }
        cand = warp_min_i(cand);
        if (cand == 0x7fffffff) break;
        double v = x_literal(s_den[cand], f_hz);
        if (mode == 1) v = __dadd_rn(v, y_literal(s_b[cand], f_hz));
        if (v > 1.0) { jstar = cand; v_jstar = v; break; }
        if (v == 1.0) any_eq1 = true;
        start = cand + 1;
      }
      if (jstar != 0x7fffffff && jstar > 0) {
        double amax = -CUDART_INF, smax = 0.0;
        for (int k = lane; k < jstar; k += 32) {
          amax = fmax(amax, crit[k]);
          smax = fmax(smax, fabs(s_den[k] * kx) + fabs(s_b[k] * ky));
        }
        amax = warp_max(amax);
        const double thr = amax - 2.0 * kScreenTol * warp_max(smax);
        for (int k = lane; k < jstar; k += 32) {
          if (crit[k] >= thr) {
            double v = x_literal(s_den[k], f_hz);
            if (mode == 1) v = __dadd_rn(v, y_literal(s_b[k], f_hz));
            M = fmax(M, v);
          }
        }
        M = warp_max(M);                                   // cummax[jstar-1], literal
      }
  } else {
      int first_gt = 0x7fffffff;
      bool any_ge = false;
      for (int k = lane; k < nt; k += 32) {
        double v = x_literal(s_den[k], f_hz);
        if (mode == 1) v = __dadd_rn(v, y_literal(s_b[k], f_hz));
        crit[k] = v;
        has_nan |= isnan(v);
        any_ge |= (v >= 1.0);
        if (v > 1.0) first_gt = min(first_gt, k);
      }
      jstar = warp_min_i(first_gt);
      has_nan = __any_sync(0xffffffffu, has_nan);
      any_eq1 = __any_sync(0xffffffffu, any_ge);           // with jstar == none this means max == 1.0
      __syncwarp();
      if (jstar != 0x7fffffff && jstar > 0) {
        double pm = -CUDART_INF;
        for (int k = lane; k < jstar; k += 32) pm = fmax(pm, crit[k]);
        M = warp_max(pm);
        v_jstar = crit[jstar];
      }
    }
  RowScan out;
  out.jstar = jstar;
  out.v_jstar = v_jstar;
  out.M = M;
  out.any_eq1 = any_eq1;
  out.has_nan = has_nan;
  return out;
}

// numpy's single-node np.interp has no NaN test: interp(NaN, [x0], [f0]) == f0.  A dead row of a one-level
// profile therefore still sees finite den/bmag/bpsi, every dh is NaN except the final 1e-6 (lib:416), and
// the reference returns alt_min + mu'(level 0) * 1e-6.
__device__ __noinline__ double dead_row_single_level(int mode, bool iso, double den0, double b0, double psi0,
                                                     double f_hz, double alt_min) {
  const double X = x_literal(den0, f_hz);
  double mup;
  if (iso) mup = iso_mup(X, nullptr);
  else if (mode == 0) mup = ah_literal<0>(X, y_literal(b0, f_hz), psi0, nullptr);
  else mup = ah_literal<1>(X, y_literal(b0, f_hz), psi0, nullptr);
  const double term = mup * kBackoff;
  return (term == term && term != 0.0) ? term + alt_min : CUDART_NAN;
}

// Rows that reflect at or below the first level (h_c <= alt0, e.g. every X-mode row below the gyrofrequency): np.interp
// clamps every grid point to level 0, so mu' is ONE number and the row's nansum is mu' * sum(dh_i) (lib:288).
// mu' at level 0 on the evaluation path the tile kernels would take (const_mup_sum).
__device__ __noinline__ double level0_mup(int mode, bool iso, bool literal, double den0, double b0, double psi0,
                                          double f_hz) {
  const double X = x_literal(den0, f_hz);
  if (iso) return iso_mup(X, nullptr);
  const double Y = y_literal(b0, f_hz);
  if (literal) return (mode == 0) ? ah_literal<0>(X, Y, psi0, nullptr) : ah_literal<1>(X, Y, psi0, nullptr);
  double sn, cs;
  sincos(__dmul_rn(psi0, kDeg2Rad), &sn, &cs);
  return (mode == 0) ? ah_fast<0>(X, Y, sn, cs, nullptr) : ah_fast<1>(X, Y, sn, cs, nullptr);
}
// The whole row in closed form (n_points >= 2, mu' finite or NaN): the weights telescope, sum(dh_i) = h_{n-1} - h_0 +
// 1e-6 (lib:415-416) with h_0 = alt0 (m_0 = 0) and h_{n-1} = span + alt0 (m_{n-1} = 1); a NaN mu' makes every term NaN,
// nansum drops them all and the empty sum becomes NaN (lib:290).
__device__ __forceinline__ double clamped_row_vh(double mup0, double span, double alt0, double alt_min) {
  if (mup0 != mup0) return CUDART_NAN;
  const double h_b = __dadd_rn(span, alt0);                         // lib:413 at m = 1
  const double total = fma(mup0, kBackoff, mup0 * __dsub_rn(h_b, alt0));
  return (total == 0.0) ? CUDART_NAN : total + alt_min;             // lib:290, lib:292
}

// Queue of the rows the tile kernel still has to do: one atomic per warp, entries of the warp's rows adjacent.
__device__ __forceinline__ void append_live_rows(const VfoParams& p, bool live, int64_t lrow, double span) {
  const unsigned mask = __activemask();
  const unsigned votes = __ballot_sync(mask, live);
  if (votes == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(votes) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(p.live_count, (unsigned)__popc(votes));
  base = __shfl_sync(mask, base, leader);
  if (live) {
    LiveRow e;
    e.row = (int)lrow;
    e.pad = 0;
    e.span = span;
    p.live_list[base + __popc(votes & ((1u << lane) - 1u))] = e;
  }
}

// `item` = (profile index inside the launch) * chunks + chunk of sounding frequencies.
// Solo mode (p.k1_solo): the CTA belongs to ONE row (item = row index inside the launch); warp 0 scans it and
// the outcome is also left in shared memory (*s_rec, *s_span) for the tile code that follows in the same CTA.
__device__ __forceinline__ void rows_body(const VfoParams& p, const int mode, const int64_t item, double* smem,
                                          BlockScratch& sc, ProfileRecord* s_rec, double* s_span) {
  const int A = p.n_alt;
  // levels staged in shared memory -- or, for profiles with more levels than it holds (p.levels_in_global; the host
  // then always asks for the thread-per-frequency mapping, which needs no per-warp scratch), read in place
  const bool glob = p.levels_in_global != 0;
  double* w_den = smem;
  double* w_alt = w_den + A;
  double* w_b = w_alt + A;
  double* w_psi = w_b + A;
  double* s_crit = w_psi + A;    // [kRowsPerCta][A]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rows_per_cta = p.k1_solo ? 1 : ((p.k1_lane_mode || p.levels_in_global) ? kThreads : kRowsPerCta * p.rows_per_warp);
  const int chunks = (p.n_freq + rows_per_cta - 1) / rows_per_cta;
  const int64_t lprof = item / chunks;                  // profile index inside this launch
  const int g = (int)(item % chunks);
  const int64_t prof = p.profile_offset + lprof;

  const double* g_den = p.den + prof * A;
  const double* g_b = p.bmag + prof * A;
  const double* g_psi = p.bpsi + prof * A;
  const double* g_alt = p.alt + prof * p.alt_stride;
  const double* s_den = glob ? g_den : w_den;
  const double* s_alt = glob ? g_alt : w_alt;
  const double* s_b = glob ? g_b : w_b;
  const double* s_psi = glob ? g_psi : w_psi;

  PRHF_TRACE_K1(0);
#ifdef PRHF_TRACE
  if (p.trace_k1 && threadIdx.x == 0) p.trace_k1[(size_t)blockIdx.x * 8 + 6] = trace_globaltimer();
#endif
  // ---- stage the whole profile with all loads in flight at once (one DRAM latency), then
  //      argmax(den) (lib:371) and min(alt) (lib:507) ----
  double best_v = -CUDART_INF;
  int best_i = 0x7fffffff;
  double amin = CUDART_INF;
  {
    // the first kPre levels of every thread are loaded into registers before anything consumes them: with a
    // plain loop the argmax test made every trip wait for its own DRAM round trip (3 in a row for 620 levels)
    constexpr int kPre = 3;
    double d[kPre], a[kPre], b[kPre], ps[kPre];
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int k = tid + u * kThreads;
      if (k < A) { d[u] = g_den[k]; a[u] = g_alt[k]; b[u] = g_b[k]; ps[u] = g_psi[k]; }
    }
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int k = tid + u * kThreads;
      if (k < A) {
        if (!glob) {
          w_den[k] = d[u];
          w_alt[k] = a[u];
          w_b[k] = b[u];
          w_psi[k] = ps[u];
        }
        if (arg_precedes(d[u], k, best_v, best_i)) { best_v = d[u]; best_i = k; }
        amin = fmin(amin, a[u]);
      }
    }
    for (int k = tid + kPre * kThreads; k < A; k += kThreads) {
      const double dk = g_den[k], ak = g_alt[k], bk = g_b[k], pk = g_psi[k];
      if (!glob) {
        w_den[k] = dk;
        w_alt[k] = ak;
        w_b[k] = bk;
        w_psi[k] = pk;
      }
      if (arg_precedes(dk, k, best_v, best_i)) { best_v = dk; best_i = k; }
      amin = fmin(amin, ak);
    }
  }
  PRHF_TRACE_K1(1);
  __shared__ long long s_key[kThreads / 32];
  __shared__ int s_ri[kThreads / 32];
  __shared__ double s_ra[kThreads / 32], s_rb[kThreads / 32];
  __syncthreads();                                      // the staged profile is visible to every warp
  const ProfileReduce1 r1 = block_argmax_min(best_v, best_i, amin, s_key, s_ri, s_ra);
  const int nt = r1.nt;                                 // truncated length = index of the peak
  const double alt_min = r1.alt_min;
  PRHF_TRACE_K1(2);

  // ---- node checks over [0, nt): negative density, non-finite values, angle steps, grid uniformity, max|B| ----
  int chk = 0;                                          // bit 0 negative density, 1 general path, 2 non-uniform grid
  double bmax = 0.0, step_max = 0.0;
  const double alt0 = s_alt[0];
  const double mean_step = (nt > 1) ? (s_alt[nt - 1] - alt0) * rcp_fast((double)(nt - 1)) : 1.0;
  for (int k = tid; k < nt; k += kThreads) {
    const double d = s_den[k];
    const double b = s_b[k];
    const double ps = s_psi[k];
    const double a = s_alt[k];
    if (d < 0.0) chk |= 1;
    // finite <=> (x - x) == 0
    if (!(((d - d) + (b - b)) + ((ps - ps) + (a - a)) == 0.0)) chk |= 2;
    if (p.n_points > 1 || k == 0) bmax = fmax(bmax, fabs(b));   // n_points == 1: the only grid point is level 0
    if (!(fabs(a - fma((double)k, mean_step, alt0)) <= 2e-15 * fmax(fabs(a), mean_step))) chk |= 4;   // ~8 ulp
    if (!(a > 0.0)) chk |= 2;                           // the fast paths compare altitudes as integers
    if (k + 1 < nt) {
      if (!(__dsub_rn(s_alt[k + 1], a) > 0.0)) chk |= 2 | 8;   // not strictly increasing: general path, np.interp range tests
      const double step = fabs(__dsub_rn(s_psi[k + 1], ps)) * kDeg2Rad;
      if (!(step <= kMaxRotateStep)) chk |= 2;
      step_max = fmax(step_max, step);
      // the quadratic form of the small-rotation path (stage_nodes, E-space) drops (dB/B) (d psi)^2 / 2 per level: it
      // must stay at the 1e-11 the rotation's own truncation (d psi)^3 / 6 is held to -- next to the reflection level
      // mu^2 cancels to 1e-8 of its operands and a SYSTEMATIC relative error of the field terms is amplified
      // accordingly (a bound of 1e-3 on dB/B alone let a random test profile through at 2.2e-9)
      if (!(fabs(__dsub_rn(s_b[k + 1], b)) * (step * step) <= 2e-11 * fabs(b))) chk |= 16;
    }
  }
  PRHF_TRACE_X(0);
  __syncthreads();                                      // s_key/s_ri/s_ra are free again
  const ProfileReduce2 r2 = block_max2_or(bmax, step_max, chk, s_ra, s_rb, s_ri);
  PRHF_TRACE_X(1);
  bmax = r2.bmax;
  step_max = r2.step_max;
  const bool any_neg = (r2.flags & 1) != 0;
  const bool any_general = (r2.flags & 2) != 0;
  const bool any_nonuniform = (r2.flags & 4) != 0;
  const int status = (nt == 0) ? 2 : (any_neg ? 1 : 0);  // lib:399 IndexError / lib:93-94 ValueError

  // Unmagnetised switch (lib:201), decided per profile from the node values: isotropic iff
  // g_p * max|B| / min|f| < 1e-12 over the profile's frequencies.  (The reference takes nanmax|Y|
  // over the regridded [F x N] array of one call; the two differ only for |B| ~ 1e-17 T.)
  bool iso = false;
  if (status == 0 && bmax < 1e-9) {
    double fmin_abs = CUDART_INF;
    for (int k = tid; k < p.n_freq; k += kThreads) {
      const double f = fabs(__dmul_rn(p.freq[prof * p.freq_stride + k], p.freq_scale));
      if (f > 0.0) fmin_abs = fmin(fmin_abs, f);
    }
    fmin_abs = block_min(fmin_abs, sc);
    iso = (bmax == 0.0) || (__ddiv_rn(__dmul_rn(kGp, bmax), fmin_abs) < kYTol);
  }

  if ((g == 0 || s_rec != nullptr) && tid == 0) {
    ProfileRecord rec;
    rec.nt = nt;
    rec.flags = (iso ? kFlagIso : 0) | (any_general ? kFlagGeneral : 0) | (status ? kFlagFailed : 0) |
                (step_max == 0.0 ? kFlagPsiConst : 0) | ((step_max <= kSmallRotateStep && !(r2.flags & 16)) ? kFlagPsiSmall : 0) |
                (any_nonuniform ? 0 : kFlagUniformAlt) | ((r2.flags & 8) ? kFlagAltUnsorted : 0);
    rec.alt_min = alt_min;
    rec.inv_dalt = (nt > 1) ? (double)(nt - 1) * rcp_fast(s_alt[nt - 1] - alt0) : 0.0;
    rec.alt0 = alt0;
    sincos(s_psi[0] * kDeg2Rad, &rec.sn0, &rec.cs0);
    rec.pad[0] = rec.pad[1] = 0.0;
    if (s_rec) *s_rec = rec;
    if (g == 0) {
      p.prof_rec[lprof] = rec;
      if (p.status) p.status[prof] = status;
    }
  }
  PRHF_TRACE_K1(3);
  // ---- lane-per-row scan (large batches): thread t takes sounding frequency g * 256 + t ----
  // Same decisions as the warp-per-row scan below, organised so that the loop over the profile levels is
  // uniform across the warp (shared-memory broadcasts, no divergence) and the literal evaluations happen
  // in lock-step: pass A screens every level and remembers, per row, the first level that may reach 1
  // (cand) and the largest screen value below it (kmax) with its runner-up; pass B evaluates those two
  // levels literally.  Rows where the screen cannot separate the candidates (values within kScreenTol of a
  // decision, or a literal candidate that turns out <= 1) fall back to a sequential literal scan.
  if (p.k1_lane_mode || glob) {
    const int r = g * kThreads + tid;
    if (r >= p.n_freq) return;
    const int64_t out_idx = prof * p.n_freq + r;
    const int64_t lrow = lprof * p.n_freq + r;
    if (status != 0) {
      p.vh[out_idx] = CUDART_NAN;
      p.row_span[lrow] = CUDART_NAN;
      if (p.row_hc) p.row_hc[lrow] = CUDART_NAN;
      return;
    }
    // (queued mode appends with a vote among the threads that reach the append together: __activemask)
    const double f_hz = __dmul_rn(p.freq[prof * p.freq_stride + r], p.freq_scale);   // lib:491
    double kx, ky;
    row_scales(f_hz, &kx, &ky);
    if (mode != 1) ky = 0.0;
    bool slow = any_general || !(isfinite(kx) && isfinite(ky) && kx > 0.0);
    int jstar = 0x7fffffff;
    bool any_eq1 = false, has_nan = false;
    double v_jstar = 0.0, M = -CUDART_INF;
    if (!slow) {
      int cand = -1, kmax = -1;
      double amax = -CUDART_INF, a2 = -CUDART_INF, smax = 0.0;
      for (int k = 0; k < nt; ++k) {
        const double xk = s_den[k] * kx, yk = s_b[k] * ky;
        const double a = xk + yk, sabs = fabs(xk) + fabs(yk);
        if (cand < 0) {
          if (a >= 1.0 - kScreenTol * sabs) {
            cand = k;
          } else {
            if (a > amax) { a2 = amax; amax = a; kmax = k; }
            else if (a > a2) a2 = a;
            smax = fmax(smax, sabs);
          }
        }
        if ((k & 15) == 15 && __all_sync(__activemask(), cand >= 0)) break;
      }
      if (cand >= 0) {
        double v = x_literal(s_den[cand], f_hz);
        if (mode == 1) v = __dadd_rn(v, y_literal(s_b[cand], f_hz));
        if (v > 1.0) {
          jstar = cand;
          v_jstar = v;
          if (cand > 0) {
            if (a2 < amax - 2.0 * kScreenTol * smax) {    // one level clearly holds the running max
              M = x_literal(s_den[kmax], f_hz);
              if (mode == 1) M = __dadd_rn(M, y_literal(s_b[kmax], f_hz));
            } else {
              slow = true;
            }
          }
        } else {
          slow = true;                                    // the literal value did not cross: scan on literally
        }
      }
    }
    if (slow) {                                           // sequential literal scan (lib:380-399 as written)
      jstar = 0x7fffffff;
      M = -CUDART_INF;
      double run = 0.0;
      for (int k = 0; k < nt; ++k) {
        double v = x_literal(s_den[k], f_hz);
        if (mode == 1) v = __dadd_rn(v, y_literal(s_b[k], f_hz));
        has_nan |= isnan(v);
        if (jstar == 0x7fffffff) {
          if (v > 1.0) { jstar = k; v_jstar = v; M = run; }
          else { run = (k == 0) ? v : fmax(run, v); any_eq1 |= (v == 1.0); }
        }
      }
    }
    const bool dead = has_nan || (jstar == 0x7fffffff && !any_eq1);
    if (dead) {                                           // valid == False (lib:399) -> NaN (lib:407)
      const double res = (nt == 1) ? dead_row_single_level(mode, iso, s_den[0], s_b[0], s_psi[0], f_hz, alt_min)
                                   : CUDART_NAN;
      p.vh[out_idx] = res;
      p.row_span[lrow] = CUDART_NAN;
      if (p.row_hc) p.row_hc[lrow] = CUDART_NAN;
      return;
    }
    double hcrit;
    if (jstar == 0 || nt == 1) {
      hcrit = s_alt[0];
    } else if (jstar == 0x7fffffff) {
      hcrit = s_alt[nt - 1];
    } else {
      const int j = jstar - 1;
      if (M == 1.0) {
        hcrit = s_alt[j];
      } else {
        const double slope = __ddiv_rn(__dsub_rn(s_alt[j + 1], s_alt[j]), __dsub_rn(v_jstar, M));
        hcrit = __dadd_rn(__dmul_rn(slope, __dsub_rn(1.0, M)), s_alt[j]);
      }
    }
    double span = __dsub_rn(__dsub_rn(hcrit, kBackoff), s_alt[0]);        // lib:407, lib:413
    if (p.row_hc) p.row_hc[lrow] = __dsub_rn(hcrit, kBackoff);
    // Queued mode: rows clamped to the first level are finished here in closed form (a fifth of the reflecting rows of a
    // global grid; in the tile kernel each would hold a CTA slot for its prologue), so that every queued row costs the
    // same n_points grid points and the tile kernel's static stride over the queue stays balanced.  (An infinite mu' --
    // mu == 0 exactly at level 0 -- keeps the term-by-term loop of the tile kernel: inf and -inf terms make NaN.)
    if (p.live_count && p.k1_finish_clamped && (!(span > 0.0) || nt == 1) && !(r2.flags & 8) && p.n_points >= 2) {
      const double mup0 = level0_mup(mode, iso, p.k1_finish_clamped == 2, s_den[0], s_b[0], s_psi[0], f_hz);
      if (!isinf(mup0)) {
        p.vh[out_idx] = clamped_row_vh(mup0, span, s_alt[0], alt_min);
        span = CUDART_NAN;                                                 // finished
      }
    }
    p.row_span[lrow] = span;
    if (p.live_count) append_live_rows(p, span == span, lrow, span);
    return;
  }

  // ---- solo mode: the CTA owns ONE row, so all 256 threads share its scan ----
  // Every thread screens its own levels (stride 256); the first level that may reach 1 (cand) comes from one
  // block reduction; the thread owning cand and the threads owning the 255 levels below it evaluate theirs
  // literally in lock-step (one sqrt/divide latency for the whole block); a second reduction yields the
  // literal running max M below cand and the largest screen value further down, which must stay clearly
  // below M.  Anything unusual (literal value at cand not above 1, a deep level competing with M, non-finite
  // inputs) leaves solo_quick false and warp 0 runs the general scan below.
  __shared__ int s_solo_quick, s_solo_cand;
  __shared__ double s_solo_vc, s_solo_M;
  if (p.k1_solo) {
    if (tid == 0) s_solo_quick = 0;
    const double f_hz0 = __dmul_rn(p.freq[prof * p.freq_stride + g], p.freq_scale);
    double kx0, ky0;
    row_scales(f_hz0, &kx0, &ky0);
    if (mode != 1) ky0 = 0.0;
    const bool screen0 = status == 0 && !any_general && isfinite(kx0) && isfinite(ky0) && kx0 > 0.0;
    if (screen0) {                                        // block-uniform
      int my_cand = 0x7fffffff;
      for (int k = tid; k < nt; k += kThreads) {
        const double xk = s_den[k] * kx0, yk = s_b[k] * ky0;
        if (my_cand == 0x7fffffff && xk + yk >= 1.0 - kScreenTol * (fabs(xk) + fabs(yk))) my_cand = k;
      }
      PRHF_TRACE_X(3);
      my_cand = warp_min_i(my_cand);
      __syncthreads();
      if (lane == 0) s_ri[wid] = my_cand;
      __syncthreads();
      int cand = s_ri[0];
#pragma unroll
      for (int k = 1; k < kThreads / 32; ++k) cand = min(cand, s_ri[k]);
      PRHF_TRACE_X(4);
      if (cand == 0x7fffffff) {
        if (tid == 0) { s_solo_cand = cand; s_solo_quick = 1; }   // no level can reach 1: no reflection
      } else {
        // own level inside the window (cand - 256, cand]
        const int kw = cand - ((cand - tid) & (kThreads - 1));
        double v = -CUDART_INF;
        if (kw >= 0) {
          v = x_literal(s_den[kw], f_hz0);
          if (mode == 1) v = __dadd_rn(v, y_literal(s_b[kw], f_hz0));
        }
        // screen values (plus their error bound) of own levels below the window
        double deep = -CUDART_INF;
        for (int k = kw - kThreads; k >= 0; k -= kThreads) {
          const double xk = s_den[k] * kx0, yk = s_b[k] * ky0;
          deep = fmax(deep, xk + yk + kScreenTol * (fabs(xk) + fabs(yk)));
        }
        PRHF_TRACE_X(5);
        double below = (kw >= 0 && kw < cand) ? v : -CUDART_INF;
        below = warp_max(below);
        deep = warp_max(deep);
        PRHF_TRACE_X(9);
        __syncthreads();
        if (lane == 0) { s_ra[wid] = below; s_rb[wid] = deep; }
        if (kw == cand) s_solo_vc = v;
        __syncthreads();
        if (tid == 0) {
          double Mq = s_ra[0], dq = s_rb[0];
#pragma unroll
          for (int k = 1; k < kThreads / 32; ++k) { Mq = fmax(Mq, s_ra[k]); dq = fmax(dq, s_rb[k]); }
          s_solo_cand = cand;
          s_solo_M = Mq;
          s_solo_quick = (s_solo_vc > 1.0 && (cand == 0 || dq < Mq)) ? 1 : 0;
        }
      }
    }
    __syncthreads();
    PRHF_TRACE_X(6);
  }

  // ---- one warp per sounding frequency, p.rows_per_warp frequencies per warp ----
  double* crit = s_crit + (size_t)wid * A;
  for (int rr = 0; rr < p.rows_per_warp; ++rr) {
    const int r = p.k1_solo ? g : (g * p.rows_per_warp + rr) * kRowsPerCta + wid;
    if (r >= p.n_freq || (p.k1_solo && (wid != 0 || rr != 0))) break;
    const int64_t out_idx = prof * p.n_freq + r;
    const int64_t lrow = lprof * p.n_freq + r;
    if (status != 0) {
      if (lane == 0) {
        p.vh[out_idx] = CUDART_NAN;
        p.row_span[lrow] = CUDART_NAN;
        if (p.row_hc) p.row_hc[lrow] = CUDART_NAN;
      }
      continue;
    }
    const double f_hz = __dmul_rn(p.freq[prof * p.freq_stride + r], p.freq_scale);   // lib:491

    // Critical curve at the nodes (lib:380-399): v_k = X_k (O) or X_k + Y_k (X) in the reference's rounding
    // order.  The reference needs: does any v_k reach 1 (validity), the first k with v_k > 1 (jstar), the
    // running max just below it (M) and v_jstar.  Evaluating 7 correctly rounded operations per node and
    // frequency costs more than the grid loop when n_points is small, so the nodes are first screened with
    // a_k = den_k * (cp^2/f^2) + b_k * (g_p/f), whose distance from the literal value is bounded by
    // 13 ulp * (|X_k| + |Y_k|); only nodes whose screen value is within kScreenTol of a decision are
    // evaluated literally, so every decision and every value that enters h_c is the literal one.
    double kx, ky;
    row_scales(f_hz, &kx, &ky);
    if (mode != 1) ky = 0.0;
    const bool screen = !any_general && isfinite(kx) && isfinite(ky) && kx > 0.0;
    int jstar = 0x7fffffff;
    bool any_eq1 = false, has_nan = false;
    double v_jstar = 0.0, M = -CUDART_INF;
    PRHF_TRACE_X(7);
    if (p.k1_solo && s_solo_quick) {
      jstar = s_solo_cand;
      v_jstar = s_solo_vc;
      M = s_solo_M;
    } else {
      const RowScan rs = row_scan_general(screen, mode, nt, lane, f_hz, kx, ky, s_den, s_b, crit);
      jstar = rs.jstar;
      v_jstar = rs.v_jstar;
      M = rs.M;
      any_eq1 = rs.any_eq1;
      has_nan = rs.has_nan;
    }
    const bool dead = has_nan || (jstar == 0x7fffffff && !any_eq1);
    if (dead) {                                             // valid == False (lib:399) -> NaN (lib:407)
      if (lane == 0) {
        p.vh[out_idx] = (nt == 1) ? dead_row_single_level(mode, iso, s_den[0], s_b[0], s_psi[0], f_hz, alt_min)
                                  : CUDART_NAN;
        p.row_span[lrow] = CUDART_NAN;
        if (p.row_hc) p.row_hc[lrow] = CUDART_NAN;
      }
      __syncwarp();
      continue;
    }
    // ---- reflection height: np.interp(1.0, running max, alt) (lib:403-404) ----
    double hcrit;
    if (jstar == 0 || nt == 1) {
      hcrit = s_alt[0];                                     // 1.0 < fcrit[0]: np.interp clamps left
    } else if (jstar == 0x7fffffff) {
      hcrit = s_alt[nt - 1];                                // running max ends exactly at 1.0
    } else {
      const int j = jstar - 1;
      if (M == 1.0) {
        hcrit = s_alt[j];
      } else {
        const double slope = __ddiv_rn(__dsub_rn(s_alt[j + 1], s_alt[j]), __dsub_rn(v_jstar, M));
        hcrit = __dadd_rn(__dmul_rn(slope, __dsub_rn(1.0, M)), s_alt[j]);
      }
    }
    PRHF_TRACE_X(8);
    if (lane == 0) {
      const double hc = __dsub_rn(hcrit, kBackoff);         // lib:407
      const double span = __dsub_rn(hc, s_alt[0]);          // lib:413 (h_c - aalt[0])
      p.row_span[lrow] = span;
      if (p.row_hc) p.row_hc[lrow] = hc;
      if (s_span) *s_span = span;
      if (p.live_count) {                                   // planned mode: compact list of rows that reflect
        LiveRow e;
        e.row = (int)lrow;
        e.pad = 0;
        e.span = span;
        p.live_list[atomicAdd(p.live_count, 1u)] = e;
      }
    }
    __syncwarp();
  }
  PRHF_TRACE_K1(4);
#ifdef PRHF_TRACE
  if (p.trace_k1 && threadIdx.x == 0) p.trace_k1[(size_t)blockIdx.x * 8 + 7] = trace_globaltimer();
#endif
}

#ifndef PRHF_ROWS_MINB
#define PRHF_ROWS_MINB 4                                    // 64 registers: 0.43 -> 0.29 ms per 4096-profile launch (3: 0.33)
#endif
__global__ void __launch_bounds__(kThreads, PRHF_ROWS_MINB) vfo_rows_kernel(const VfoParams p, const int mode) {
  extern __shared__ __align__(16) double smem[];
  __shared__ BlockScratch sc;
  // Programmatic dependent launch: the tile kernel may be scheduled as soon as every CTA of this grid has
  // started; it blocks in griddepcontrol.wait until this grid has completed and its writes are visible.
  asm volatile("griddepcontrol.launch_dependents;");
  rows_body(p, mode, blockIdx.x, smem, sc, nullptr, nullptr);
}

// ==========================================================================================
// K2: grid points
// ==========================================================================================
// Node table entry in shared memory (one per staged profile level): 64 bytes, read as 4 x LDS.128.
// Fast paths: density and field are pre-multiplied by the row's cp^2/f^2 and g_p/f, so the per-point
// interpolation yields X and Y directly.
struct __align__(16) Node {
  double alt, x;       // level altitude, X at the level            (general paths: density)
  double sx, y;        // slope of X, Y at the level                (general paths: density slope, field)
  double sy, srad;     // slope of Y, field-angle slope in rad/km   (general paths: field slope, angle slope deg/km)
  double sn, cs;       // sin / cos of the field angle at the level (general paths: angle in degrees, unused)
};

struct RowConst {
  double f_hz;      // lib:491
  double alt0;      // aalt[0]
  double span;      // h_c - aalt[0]  (lib:413)
  double inv_dalt;  // (nt-1)/(alt[nt-1]-alt[0]): bracket guess for (near-)uniform altitude grids
  int nt;           // truncated length (= argmax(den))
  int jlo, jhi;     // node window staged for this tile (absolute indices)
  int lane0;        // index of this thread inside the group that shares the row (CTA: tid, warp: lane)
  int group;        // threads in that group (kTileThreads or 32)
  double kx, ky;    // cp^2/f^2 and g_p/f, applied per point when the staged nodes are not pre-scaled
  const double* g_alt;   // the profile's raw altitude / density levels: O-mode points next to the reflection level
  const double* g_den;   // re-evaluate X in numpy's operation order (near_reflection_tail)
  int unsorted;          // altitudes below the peak are not strictly increasing (kFlagAltUnsorted)
};

// np.interp on an axis that is not increasing.  What numpy does is decided by the two range tests that precede its
// search (binary_search_with_guess: key > arr[len-1] -> right fill value, else key < arr[0] -> left fill value); on a
// strictly DEcreasing grid one of them always fires, so every interpolant is fp[len-1] (or fp[0] at/below the smallest
// altitude) -- e.g. Day profile, reversed arrays: 4 finite virtual heights of 174 (tests/golden/edge.npz).  Inside the
// range of a zig-zag axis numpy's result depends on the guess it carries from query to query; a plain bisection
// stands in there (documented as undefined, as numpy documents it).  `nodes` holds all levels [0, nt).
__device__ __forceinline__ int bracket_unsorted(double h, const Node* nodes, int nt) {
  if (h > nodes[nt - 1].alt) return nt - 1;
  if (h < nodes[0].alt) return -1;
  return max(bracket_in<8>(h, &nodes[0].alt, 0, nt - 1), 0);
}

// Bracket of h inside the staged window: last j in [jlo, jhi] with alt[j] <= h (jlo - 1 if below).
// `guess` is tried first (exact for uniform grids up to rounding), then its neighbours, then bisection.
__device__ __forceinline__ int find_bracket(double h, const Node* nodes, int jlo, int jhi, int guess) {
  int j = min(max(guess, jlo), jhi);
  const double* alt = &nodes[0].alt - (ptrdiff_t)jlo * 8;      // absolute-index view, stride 8 doubles
  if (h < alt[(ptrdiff_t)j * 8]) {
    if (j == jlo) return jlo - 1;
    --j;
    if (h >= alt[(ptrdiff_t)j * 8]) return j;
    return bracket_in<8>(h, alt, jlo, j - 1);
  }
  if (j == jhi || h < alt[(ptrdiff_t)(j + 1) * 8]) return j;
  ++j;
  if (j == jhi || h < alt[(ptrdiff_t)(j + 1) * 8]) return j;
  return bracket_in<8>(h, alt, j + 1, jhi);
}

// Coordinate the fast paths stage their nodes in (stage_nodes)
enum : int { kSpaceAlt = 0, kSpaceM = 1, kSpaceE = 2 };

// Evaluation paths of the grid loop
enum : int {
  kPathFast0 = 0,    // restructured arithmetic, field angle constant with height (no rotation)
  kPathFastS = 1,    // ... angle steps <= 4e-4 rad per level: second-order rotation
  kPathFastL = 2,    // ... angle steps <= 0.05 rad per level: eighth-order rotation
  kPathGeneral = 3,  // numpy-literal interpolation + libdevice sincos + sign-safe restructured arithmetic
  kPathLiteral = 4,  // numpy-literal interpolation + reference-order arithmetic (PRHF_FLAG_LITERAL)
  kPathIso = 5       // unmagnetised branch (lib:202-206), numpy-literal interpolation
};

// mu' * dh for one grid point on the general / literal / isotropic paths (NaN -> 0 handled by the caller).
template <int MODE, int PATH>
__device__ __forceinline__ double point_term(double h, double dh, int j, const Node* nodes, const RowConst& rc) {
  double mup;
  {
    // numpy arr_interp semantics (NaN rescue, exact-node shortcut, clamping) on the staged window;
    // node fields hold raw density / field / angle[deg] and their slopes here
    const int jj = max(j, rc.jlo) - rc.jlo;
    double den, b, psi;
    if (rc.nt == 1 || j < rc.jlo) {
      den = nodes[0].x; b = nodes[0].y; psi = nodes[0].sn;           // left clamp (only when jlo == 0)
    } else if (j >= rc.nt - 1) {
      den = nodes[jj].x; b = nodes[jj].y; psi = nodes[jj].sn;
    } else {
      const Node& n0 = nodes[jj];
      const Node& n1 = nodes[jj + 1];
      if (n0.alt == h) {
        den = n0.x; b = n0.y; psi = n0.sn;
      } else {
        const double t = __dsub_rn(h, n0.alt);
        den = __dadd_rn(__dmul_rn(n0.sx, t), n0.x);
        b = __dadd_rn(__dmul_rn(n0.sy, t), n0.y);
        psi = __dadd_rn(__dmul_rn(n0.srad, t), n0.sn);
        if (isnan(den) || isnan(b) || isnan(psi)) {
          const double t1 = __dsub_rn(h, n1.alt);
          if (isnan(den)) {
            den = __dadd_rn(__dmul_rn(n0.sx, t1), n1.x);
            if (isnan(den) && n0.x == n1.x) den = n0.x;
          }
          if (isnan(b)) {
            b = __dadd_rn(__dmul_rn(n0.sy, t1), n1.y);
            if (isnan(b) && n0.y == n1.y) b = n0.y;
          }
          if (isnan(psi)) {
            psi = __dadd_rn(__dmul_rn(n0.srad, t1), n1.sn);
            if (isnan(psi) && n0.sn == n1.sn) psi = n0.sn;
          }
        }
      }
    }
    const double X = x_literal(den, rc.f_hz);                        // lib:500
    if (PATH == kPathIso) {
      mup = iso_mup(X, nullptr);
    } else {
      const double Y = y_literal(b, rc.f_hz);                        // lib:503
      if (PATH == kPathLiteral) {
        mup = ah_literal<MODE>(X, Y, psi, nullptr);
      } else {
        double sn, cs;
        sincos(__dmul_rn(psi, kDeg2Rad), &sn, &cs);
        mup = ah_fast<MODE>(X, Y, sn, cs, nullptr);
      }
    }
  }
  return mup * dh;                                                   // lib:288
}

// Grid points [i0, i1) of one row, two adjacent points per thread per iteration.
// CONST_MUP: the row reflects at or below the first level (h_c <= alt0), every grid point clamps to
// level 0 (np.interp left clamp) and mu' is one number; only the dh_i differ.
template <int MODE, int PATH, bool CONST_MUP>
__device__ __forceinline__ double tile_sum(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                           int i0, int i1, int n_points, double mup0) {
  double acc0 = 0.0, acc1 = 0.0;
  const double2* m2 = reinterpret_cast<const double2*>(m);
  for (int i = i0 + 2 * rc.lane0; i < i1; i += 2 * rc.group) {
    const double2 mm = __ldg(m2 + (i >> 1));                         // i0 is even, the table is padded
    const double mn = __ldg(m + i + 2);
    const double h0 = __dadd_rn(__dmul_rn(mm.x, rc.span), rc.alt0);  // lib:413
    const double h1 = __dadd_rn(__dmul_rn(mm.y, rc.span), rc.alt0);
    const double h2 = __dadd_rn(__dmul_rn(mn, rc.span), rc.alt0);
    const double dh0 = (i == n_points - 1) ? kBackoff : __dsub_rn(h1, h0);      // lib:415-416
    const double dh1 = (i + 1 == n_points - 1) ? kBackoff : __dsub_rn(h2, h1);
    double t0, t1;
    if (CONST_MUP) {
      t0 = mup0 * dh0;
      t1 = mup0 * dh1;
    } else {
      int j0, j1;
      if (rc.unsorted) {
        j0 = bracket_unsorted(h0, nodes, rc.nt);
        j1 = bracket_unsorted(h1, nodes, rc.nt);
      } else {
        const int g0 = __double2int_rd((h0 - rc.alt0) * rc.inv_dalt);
        j0 = find_bracket(h0, nodes, rc.jlo, rc.jhi, g0);
        j1 = find_bracket(h1, nodes, rc.jlo, rc.jhi, j0);
      }
      t0 = point_term<MODE, PATH>(h0, dh0, j0, nodes, rc);
      t1 = point_term<MODE, PATH>(h1, dh1, j1, nodes, rc);
    }
    acc0 += (t0 == t0) ? t0 : 0.0;                                   // nansum
    acc1 += (t1 == t1 && i + 1 < i1) ? t1 : 0.0;
  }
  return acc0 + acc1;
}

// ---- hot loop of the fast paths ----
// Differences from tile_sum: altitudes are positive on these paths (K1 flag), so brackets are tested on the
// IEEE bit patterns with integer compares; h_i = fma(m_i, span, alt0) (<= 0.5 ulp from the reference's
// multiply-then-add, lib:413); validity tests are integer; terms enter the sum through one FMA.
__device__ __forceinline__ int find_bracket_pos(double h, const Node* nodes, int jlo, int jhi, int guess) {
  int j = min(max(guess, jlo), jhi);
  const long long hb = __double_as_longlong(h);
  const long long* alt = reinterpret_cast<const long long*>(&nodes[0].alt) - (ptrdiff_t)jlo * 8;
  if (hb < alt[(ptrdiff_t)j * 8]) {
    if (j == jlo) return jlo;                                // (rows on this path never fall below level jlo)
    --j;
    if (hb >= alt[(ptrdiff_t)j * 8]) return j;
    return max(bracket_in<8>(h, &nodes[0].alt - (ptrdiff_t)jlo * 8, jlo, j - 1), jlo);
  }
  if (j == jhi || hb < alt[(ptrdiff_t)(j + 1) * 8]) return j;
  ++j;
  if (j == jhi || hb < alt[(ptrdiff_t)(j + 1) * 8]) return j;
  return bracket_in<8>(h, &nodes[0].alt - (ptrdiff_t)jlo * 8, j + 1, jhi);
}

// ROWSCALE: the staged nodes are shared by several rows (row-per-warp kernel) and hold density / field
// un-multiplied; the row's cp^2/f^2 and g_p/f are applied here (two more FP64 multiplies per point).
// X and the two field terms of ah_hot at one grid point, from the staged levels.
// ABS: `nodes` is the absolute-level view (staged window - jlo), indexed by j itself.
template <int PATH, bool ROWSCALE>
__device__ __forceinline__ void fast_xy_node(double h, const Node& nd, const RowConst& rc, double* X_out,
                                             double* yth_out, double* yl_out);
// E-space nodes of the constant-angle path hold intercepts and slopes (stage_nodes): one FMA per interpolant.
template <int PATH>
__device__ __forceinline__ void espace_xy_node(double E, const Node& nd, const RowConst& rc, double* X_out,
                                               double* yth_out, double* yl_out) {
  if (PATH == kPathFast0) {
    *X_out = fma(nd.x, E, nd.alt);
    *yth_out = fma(nd.y, E, nd.sx);
    *yl_out = fma(nd.srad, E, nd.sy);
  } else if (PATH == kPathFastS) {
    *X_out = fma(nd.x, E, nd.alt);
    *yth_out = fma(fma(nd.sy, E, nd.y), E, nd.sx);
    *yl_out = fma(fma(nd.cs, E, nd.sn), E, nd.srad);
  } else {
    fast_xy_node<PATH, false>(E, nd, rc, X_out, yth_out, yl_out);
  }
}
template <int PATH, bool ROWSCALE, bool ABS = false>
__device__ __forceinline__ void fast_xy(double h, int j, const Node* nodes, const RowConst& rc, double* X_out,
                                        double* yth_out, double* yl_out) {
  fast_xy_node<PATH, ROWSCALE>(h, nodes[ABS ? j : j - rc.jlo], rc, X_out, yth_out, yl_out);
}
template <int PATH, bool ROWSCALE>
__device__ __forceinline__ void fast_xy_node(double h, const Node& nd, const RowConst& rc, double* X_out,
                                             double* yth_out, double* yl_out) {
  const double t = h - nd.alt;
  double X = fma(nd.sx, t, nd.x);
  if (ROWSCALE) X *= rc.kx;
  *X_out = X;
  if (PATH == kPathFast0) {
    // node fields: y = Y sin(psi)/sqrt(2), sy its slope; srad = Y cos(psi), sn = its slope
    double yth = fma(nd.sy, t, nd.y), yl = fma(nd.sn, t, nd.srad);
    if (ROWSCALE) { yth *= rc.ky; yl *= rc.ky; }
    *yth_out = yth;
    *yl_out = yl;
    return;
  }
  double Y = fma(nd.sy, t, nd.y);
  if (ROWSCALE) Y *= rc.ky;
  double sn, cs;
  if (PATH == kPathFastS) rotate_sincos_small(nd.sn, nd.cs, nd.srad * t, &sn, &cs);
  else rotate_sincos(nd.sn, nd.cs, nd.srad * t, &sn, &cs);
  *yth_out = (Y * sn) * 0.70710678118654752;
  *yl_out = Y * cs;
}

// 0 < 1 - X < 1e-7, tested on the high word (0x3E7AD7F2 is the high word of 1e-7); see near_reflection_tail.
__device__ __forceinline__ bool near_reflection(double X) { return (unsigned)__double2hiint(1.0 - X) < 0x3E7AD7F2u; }

template <int MODE, int PATH, bool ROWSCALE, bool ABS = false>
__device__ __forceinline__ double fast_point(double h, int j, const Node* nodes, const RowConst& rc, double* mu_out,
                                             double* q_out, bool* near_out) {
  double X, yth, yl;
  fast_xy<PATH, ROWSCALE, ABS>(h, j, nodes, rc, &X, &yth, &yl);
  *near_out = (MODE == 0) && near_reflection(X);
  return ah_hot<MODE>(X, yth, yl, mu_out, q_out);
}

// X in the reference's own operation order: h = m * span + alt0 in two roundings (lib:413), np.interp on the raw levels
// (lib:424), X = (sqrt(n) cp)^2 / f^2 (lib:136).
__device__ __forceinline__ double literal_x(double mval, int j, const RowConst& rc) {
  const double h = __dadd_rn(__dmul_rn(mval, rc.span), rc.alt0);
  const double* xp = rc.g_alt;
  const int n = rc.nt;
  j = min(max(j, 0), n - 1);
  while (j > 0 && h < xp[j]) --j;                         // the fast bracket may be one off next to a level
  while (j + 1 < n && h >= xp[j + 1]) ++j;
  return x_literal(np_interp_at(h, (h < xp[0]) ? -1 : j, xp, rc.g_den, n), rc.f_hz);
}

// Bracket for altitude grids K1 flagged uniform (every level within ~8 ulp of alt0 + k * mean step, e.g. any
// np.arange / np.linspace grid): floor((h - alt0) / step), clamped to the staged window, with NO verification
// against the node altitudes.  The guess can only be off by one when h lies within ~1e-12 km of a level; the
// piecewise-linear interpolants are continuous there, so evaluating the neighbouring segment changes X, Y by
// |h - level| * |slope difference| < 1e-13 relative -- four orders below the parity tolerance -- while the
// verification cost 10 of the loop's 187 instructions per point pair.  Grids that are merely close to uniform
// take find_bracket_pos (guess, verify, binary search).
__device__ __forceinline__ int bracket_uniform(int jlo, int jhi, int guess) { return min(max(guess, jlo), jhi); }

// m-space form of the same bracket: floor(m * c1) with c1 = span * inv_dalt, taken from ONE round-down FMA onto
// 1.5 * 2^52 (the integer part of the exact product lands in the low word: no F2I, no separate multiply).  The tile's
// window [jlo, jhi] is computed with this very function from the tile's first and last multiplier, and both the
// multiplier table and x -> floor(x * c1) are monotone, so every grid point of the tile falls inside the window by
// construction: the loop needs no clamp.  (m <= 1 and c1 < nt - 1 -- the 1e-6 km back-off of lib:407 is nine orders
// above the rounding of c1 -- so the index never exceeds nt - 2 + 1 either.)
__device__ __forceinline__ int bracket_floor_m(double m, double c1) {
  return __double2loint(__fma_rd(m, c1, 6755399441055744.0));
}

// O-mode grid points with 1 - X < 1e-7 (the last few of a 20 000-point row).  mu' ~ (1 - X)^(-1/2) there and 1 - X
// goes down to 1e-9 at the last point, so ONE ulp of X is worth up to 1e-7 of the term and, on steep profiles, more
// than 1e-9 of the virtual height (conditioning; found by tools/fuzz_more.py).  The fast
// interpolant (one FMA on pre-scaled levels, h from one FMA) is a few ulp from numpy's, so the hot loop leaves these
// points out (one compare each) and this out-of-line pass adds them with X in the reference's own operation order.
// Same thread, same stride and same weights as the hot loop; `first` is the first pair it skipped a point in.
template <int PATH, bool UNIFORM, bool ROWSCALE>
__device__ __noinline__ double near_reflection_tail(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                    int first, int i1, int n_points) {
  const double c1 = rc.span * rc.inv_dalt;
  double acc = 0.0;
  for (int i = first; i < i1; i += 2 * rc.group) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = i + u;
      const double mk = __ldg(m + k), mk1 = __ldg(m + k + 1);
      const double h = fma(mk, rc.span, rc.alt0);
      const double dh = fma(mk1, rc.span, rc.alt0) - h;
      const int g = __double2int_rd(mk * c1);
      const int j = UNIFORM ? min(max(g, rc.jlo), rc.jhi) : find_bracket_pos(h, nodes, rc.jlo, rc.jhi, g);
      double X, yth, yl, mu, q;
      fast_xy<PATH, ROWSCALE>(h, j, nodes, rc, &X, &yth, &yl);
      if (!near_reflection(X) || k >= i1 || k == n_points - 1) continue;   // (the last point is added by the caller)
      const double p = ah_hot<0>(literal_x(mk, j, rc), yth, yl, &mu, &q);
      acc = fma(keep_term(p, q) ? p : 0.0, dh, acc);
    }
  }
  return acc;
}

template <int MODE, int PATH, bool UNIFORM, bool ROWSCALE>
__device__ __forceinline__ double tile_sum_fast(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                int i0, int i1, int n_points) {
  double acc0 = 0.0, acc1 = 0.0;
  int first_near = 0x7fffffff;
  const int il_row = n_points - 1;
  const double2* m2 = reinterpret_cast<const double2*>(m);
  const double c1 = rc.span * rc.inv_dalt;                           // bracket guess = floor(m_i * c1)
  // The tile kernels (long segments) read the multiplier table one iteration ahead, so that the L2 / L1 latency of
  // the load overlaps the ~95 FP64 instructions of the current pair of points (+3.9 % on batches, measured:
  // profiles/sweep_mpref_r01.log).  The row-per-warp kernel (<= 4096 points per row, 2-3 iterations per thread)
  // loses from the extra prologue load and keeps the plain form.
  constexpr bool kAhead = !ROWSCALE;
  double2 mm_next = make_double2(1.0, 1.0);
  double mn_next = 1.0;
  const double2* pm2 = m2 + ((i0 >> 1) + rc.lane0);                  // running pointers of the read-ahead
  const double* pm1 = m + (i0 + 2 * rc.lane0 + 2);
  if (kAhead) {
    const int ip = i0 + 2 * rc.lane0;
    if (ip < i1) { mm_next = __ldg(pm2); mn_next = __ldg(pm1); }
  }
  for (int i = i0 + 2 * rc.lane0; i < i1; i += 2 * rc.group) {
    double2 mm;
    double mn;
    if (kAhead) {
      mm = mm_next;
      mn = mn_next;
      pm2 += rc.group;
      pm1 += 2 * rc.group;
      if (i + 2 * rc.group < i1) { mm_next = __ldg(pm2); mn_next = __ldg(pm1); }
    } else {
      mm = __ldg(m2 + (i >> 1));                                     // i0 is even, the table is padded
      mn = __ldg(m + i + 2);
    }
    const double h0 = fma(mm.x, rc.span, rc.alt0);                   // lib:413
    const double h1 = fma(mm.y, rc.span, rc.alt0);
    const double h2 = fma(mn, rc.span, rc.alt0);
    const double dh0 = h1 - h0, dh1 = h2 - h1;                       // lib:415; 0 for the row's last point (padded table)
    int j0, j1;
    if (UNIFORM) {
      j0 = bracket_uniform(rc.jlo, rc.jhi, __double2int_rd(mm.x * c1));
      j1 = bracket_uniform(rc.jlo, rc.jhi, __double2int_rd(mm.y * c1));
    } else {
      j0 = find_bracket_pos(h0, nodes, rc.jlo, rc.jhi, __double2int_rd(mm.x * c1));
      j1 = find_bracket_pos(h1, nodes, rc.jlo, rc.jhi, j0);
    }
    double mu0, mu1, q0, q1;
    bool near0, near1;
    const double p0 = fast_point<MODE, PATH, ROWSCALE>(h0, j0, nodes, rc, &mu0, &q0, &near0);
    const double p1 = fast_point<MODE, PATH, ROWSCALE>(h1, j1, nodes, rc, &mu1, &q1, &near1);
    if (MODE == 0) {
      // the row's very last point carries weight 0 in this loop and is evaluated after it: it must not trigger the tail
      near0 = near0 && (i != il_row);
      near1 = near1 && (i + 1 != il_row);
      if (near0 || near1) first_near = min(first_near, i);   // left to near_reflection_tail
    }
    acc0 = fma((keep_term(p0, q0) && !near0) ? p0 : 0.0, dh0, acc0);        // nansum (lib:288)
    // (when n_points is odd the last pair's second point is a pad entry: h == h_c, weight h(pad) - h(pad) == 0)
    acc1 = fma((keep_term(p1, q1) && !near1) ? p1 : 0.0, dh1, acc1);
  }
  if (MODE == 0 && first_near != 0x7fffffff)
    acc0 += near_reflection_tail<PATH, UNIFORM, ROWSCALE>(nodes, rc, m, first_near, i1, n_points);
  // lib:416: the row's last grid point weighs 1e-6.  The table is padded with copies of its last entry, so that
  // point entered the loop with weight 0; its owner adds the term here instead of two selects per iteration.
  const int il = n_points - 1, ip = il & ~1;
  if (ip >= i0 && ip < i1 && ((ip - i0) >> 1) % rc.group == rc.lane0) {
    const double ml = __ldg(m + il);
    const double hl = fma(ml, rc.span, rc.alt0);
    const int g = __double2int_rd(ml * c1);
    const int jl = UNIFORM ? bracket_uniform(rc.jlo, rc.jhi, g) : find_bracket_pos(hl, nodes, rc.jlo, rc.jhi, g);
    double Xl, ythl, yll, mul, ql;
    fast_xy<PATH, ROWSCALE>(hl, jl, nodes, rc, &Xl, &ythl, &yll);
    // (coarse grids: the last point is the only one this close, and with its weight of 1e-6 km an ulp of X is worth
    //  < 1e-11 of the virtual height -- not worth an IEEE sqrt and two divisions on one lane of a 200-point row)
    if (MODE == 0 && n_points >= 1024 && near_reflection(Xl)) Xl = literal_x(ml, jl, rc);
    const double pl = ah_hot<MODE>(Xl, ythl, yll, &mul, &ql);
    acc0 = fma(keep_term(pl, ql) ? pl : 0.0, kBackoff, acc0);
  }
  return acc0 + acc1;
}

// ---- hot loop of the tile kernels, "m-space" form ----
// The tile kernels stage their nodes for ONE row, so the row's span h_c - alt0 can be folded into the staged values:
// a node stores m_j = (alt_j - alt0) / span and slopes per unit of m, the loop interpolates on t = m_i - m_j straight
// from the stretched-grid multiplier, and the weights come from the table dm_i = m_{i+1} - m_i with the sum scaled by
// span once at the end (lib:413-416: h_i = m_i span + alt0, dh_i = h_{i+1} - h_i = span dm_i up to the rounding of
// h).  Against the altitude-space loop (tile_sum_fast, still used by the row-per-warp and the global-memory kernels,
// whose nodes are shared between rows) this drops h_i, h_{i+1}, h_{i+2} and the two subtractions for dh from every
// pair of points: 78 instead of 89 FP64 instructions per pair.  The tables are padded by kMultPad entries, so the
// read one iteration ahead needs no bounds test; validity enters through one predicated DFMA (add_kept).
template <int PATH, bool UNIFORM>
__device__ __noinline__ double near_reflection_tail_m(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                      const double* __restrict__ dm, int first, int i1, int n_points) {
  const double c1 = rc.span * rc.inv_dalt;
  double acc = 0.0;
  for (int i = first; i < i1; i += 2 * rc.group) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = i + u;
      const double mk = __ldg(m + k);
      const int j = UNIFORM ? min(max(bracket_floor_m(mk, c1), rc.jlo), rc.jhi)
                            : find_bracket_pos(mk, nodes, rc.jlo, rc.jhi, __double2int_rd(mk * c1));
      double X, yth, yl, mu, q;
      fast_xy<PATH, false>(mk, j, nodes, rc, &X, &yth, &yl);
      if (!near_reflection(X) || k >= i1 || k == n_points - 1) continue;   // (the last point is added by the caller)
      const double p = ah_hot<0>(literal_x(mk, j, rc), yth, yl, &mu, &q);
      acc = fma(keep_term(p, q) ? p : 0.0, __ldg(dm + k) * rc.span, acc);
    }
  }
  return acc;
}

template <int MODE, int PATH, bool UNIFORM>
__device__ __forceinline__ double tile_sum_fast_m(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                  const double* __restrict__ dm, int i0, int i1, int n_points) {
  double acc0 = 0.0, acc1 = 0.0;                                     // in units of the span
  int first_near = 0x7fffffff;
  const int il_row = n_points - 1;
  const double c1 = rc.span * rc.inv_dalt;                           // bracket guess = floor(m_i * c1)
  const double2* m2 = reinterpret_cast<const double2*>(m);           // i0 is even, both tables are 16-byte aligned
  const double2* d2 = reinterpret_cast<const double2*>(dm);
  int ip = (i0 >> 1) + rc.lane0;                                     // index of this thread's pair of points
  const int ip_end = (i1 + 1) >> 1;
  double2 mm_next = __ldg(m2 + ip), dd_next = __ldg(d2 + ip);
  const Node* nb = nodes - rc.jlo;                                   // absolute-level view of the staged window
#pragma unroll 2
  for (; ip < ip_end; ip += rc.group) {                              // (unrolled by two: the read-ahead registers
    const int i = 2 * ip;                                            //  alternate instead of being copied)
    const double2 mm = mm_next, dd = dd_next;
    mm_next = __ldg(m2 + ip + rc.group);                             // next iteration's entries (padded tables)
    dd_next = __ldg(d2 + ip + rc.group);
    int j0, j1;
    if (UNIFORM) {
      j0 = bracket_floor_m(mm.x, c1);
      j1 = bracket_floor_m(mm.y, c1);
    } else {
      j0 = find_bracket_pos(mm.x, nodes, rc.jlo, rc.jhi, __double2int_rd(mm.x * c1));
      j1 = find_bracket_pos(mm.y, nodes, rc.jlo, rc.jhi, j0);
    }
    double mu0, mu1, q0, q1;
    bool near0, near1;
    const double p0 = fast_point<MODE, PATH, false, true>(mm.x, j0, nb, rc, &mu0, &q0, &near0);
    const double p1 = fast_point<MODE, PATH, false, true>(mm.y, j1, nb, rc, &mu1, &q1, &near1);
    if (MODE == 0) {
      // the row's very last point carries weight 0 in this loop and is evaluated after it: it must not trigger the tail
      near0 = near0 && (i != il_row);
      near1 = near1 && (i + 1 != il_row);
      if (near0 || near1) first_near = min(first_near, i);          // left to near_reflection_tail_m
      if (!near0) add_kept(acc0, p0, q0, dd.x);                     // nansum (lib:288)
      if (!near1) add_kept(acc1, p1, q1, dd.y);
    } else {
      add_kept(acc0, p0, q0, dd.x);
      // (when n_points is odd the last pair's second point is a pad entry: m == 1, weight 0)
      add_kept(acc1, p1, q1, dd.y);
    }
  }
  double acc_km = 0.0;
  if (MODE == 0 && first_near != 0x7fffffff)
    acc_km = near_reflection_tail_m<PATH, UNIFORM>(nodes, rc, m, dm, first_near, i1, n_points);
  // lib:416: the row's last grid point weighs 1e-6 km.  Its table weight is 0, so it went through the loop for nothing;
  // its owner adds the term here instead of two selects per iteration.
  const int il = n_points - 1, ipl = il & ~1;
  if (ipl >= i0 && ipl < i1 && ((ipl - i0) >> 1) % rc.group == rc.lane0) {
    const double ml = __ldg(m + il);
    const int jl = UNIFORM ? bracket_uniform(rc.jlo, rc.jhi, bracket_floor_m(ml, c1))
                           : find_bracket_pos(ml, nodes, rc.jlo, rc.jhi, __double2int_rd(ml * c1));
    double Xl, ythl, yll, mul, ql;
    fast_xy<PATH, false>(ml, jl, nodes, rc, &Xl, &ythl, &yll);
    // (coarse grids: the last point is the only one this close, and with its weight of 1e-6 km an ulp of X is worth
    //  < 1e-11 of the virtual height -- not worth an IEEE sqrt and two divisions on one lane of a 200-point row)
    if (MODE == 0 && n_points >= 1024 && near_reflection(Xl)) Xl = literal_x(ml, jl, rc);
    const double pl = ah_hot<MODE>(Xl, ythl, yll, &mul, &ql);
    acc_km = fma(keep_term(pl, ql) ? pl : 0.0, kBackoff, acc_km);
  }
  return fma(acc0 + acc1, rc.span, acc_km);
}

// ---- hot loop of the tile kernels on uniform altitude grids, "E-space" form ----
// The stretched grid is a geometric sequence in disguise: m_i = A - B E_i with E_i = exp(10 (1 - i/(n-1))) (lib:314-320),
// so E_{i+s} = E_i * exp(-10 s/(n-1)) and the weight of point i is dm_i = m_{i+1} - m_i = B (1 - exp(-10/(n-1))) E_i.
// The tile's nodes are staged in the coordinate E (stage_nodes, kSpaceE: interpolation that is linear in m is linear in
// E), each thread seeds E for its pair of points from the table once and then advances it by ONE multiplication per
// point and iteration, and the sum is accumulated with the weights E_i and scaled by e_weight * span at the end.  The
// loop therefore reads NO table: the m-space loop pulled 32 bytes per thread and iteration through L2 (5 TB/s over the
// whole GPU, every CTA walks the whole 320 KB table of a 20 000-point row) and spent a third of its warp time waiting
// for them (profiles/ncu_r02h_*: long scoreboard 2.5 of 7.8 warps) -- at the price of one more FP64 instruction per
// point.  Rounding: the seeds are the table's own exp values; k multiplications by the correctly rounded ratio drift
// by <= k * 1.1e-16 relative and the thread re-seeds every kReseed iterations, so E_i, and with it h_i and dh_i, stay
// within 1e-14 relative of the table form -- the table's dm_i = fl(m_{i+1} - m_i) itself carries 5e-9 at the top of a
// 20 000-point row, as do the reference's own dh_i = fl(h_{i+1} - h_i).
//
// Bracket: x = c1 (A - B E) + 2^-36 levels, taken from ONE round-down FMA onto 1.5 * 2^14, whose ulp is 2^-38: the
// integer part of x (x < 4096 levels) lands in bits 6..17 of the high word, i.e. (high & 0x3FFC0) is the BYTE offset
// j * sizeof(Node) of the bracketing level -- one FMA and one AND from E to the shared-memory offset, no F2I, no
// multiply, no clamp.  Resolution 3.6e-12 levels: a point is attributed to the neighbouring segment only within that
// distance of a level, where the continuous interpolants differ by < 1e-13 even on the steepest profiles (a coarser
// 2^-20 was tried first and is NOT enough: next to the reflection level mu' is singular and a point 1e-6 km off a level
// moved one row in 10^4 by up to 4e-7, profiles/ab_r02i_espace.txt).  No clamp: x > 0 because of the 2^-36 offset (m_0 is
// 0 to ~1e-16), x < nt - 1 + 1e-10 because c1 < nt - 1 by the 1e-6 km back-off of lib:407, level nt - 1 continues the
// last segment's line (stage_nodes), and the window staged for the tile is computed with the same instruction on the
// tile's first and last point, widened by one level for the drift of the recurrence.
constexpr int kReseed = 64;
constexpr double kBracketMagic = 24576.0;                   // 1.5 * 2^14: ulp 2^-38
constexpr double kBracketOffset = 1.4551915228366852e-11;   // 2^-36 levels
constexpr int kBracketMaxLevels = 4095;
struct BracketE { double kn, km; };                         // x + magic = fma(E, kn, km)
__device__ __forceinline__ BracketE bracket_e_constants(double c1) {
  BracketE b;
  b.kn = -(c1 * kStretchB);
  b.km = fma(c1, kStretchA, kBracketOffset) + kBracketMagic;
  return b;
}
// byte offset of the bracketing level: j * 64
__device__ __forceinline__ unsigned bracket_bytes_e(double E, const BracketE& b) {
  static_assert(sizeof(Node) == 64, "bracket_bytes_e yields j * 64");
  return (unsigned)__double2hiint(__fma_rd(E, b.kn, b.km)) & 0x3FFC0u;
}
__device__ __forceinline__ int bracket_floor_e(double E, const BracketE& b) { return (int)(bracket_bytes_e(E, b) >> 6); }

// O-mode grid points with 1 - X < 1e-7 that the E-space loop left out (see near_reflection_tail): same thread, same
// pairs and -- so that the "near" decision falls exactly as it fell in the loop -- the same E_i: the recurrence is
// replayed from the seed of the re-seed block that holds the first such pair (ip_first).
template <int PATH>
__device__ __noinline__ double near_reflection_tail_e(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                      const double* __restrict__ dm, const double* __restrict__ etab,
                                                      double e_ratio, int ip, int ip_first, int ip_end) {
  const BracketE br = bracket_e_constants(rc.span * rc.inv_dalt);
  const double2* e2 = reinterpret_cast<const double2*>(etab);
  const Node* nb = nodes - rc.jlo;
  double acc = 0.0;
  while (ip + kReseed * rc.group <= ip_first) ip += kReseed * rc.group;   // whole re-seed blocks before the first pair
  while (ip < ip_end) {
    const double2 seed = __ldg(e2 + ip);
    double E[2] = {seed.x, seed.y};
    const int ip_blk = min(ip_end, ip + kReseed * rc.group);
    for (; ip < ip_blk; ip += rc.group) {
      if (ip >= ip_first) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = 2 * ip + u;
          const int j = bracket_floor_e(E[u], br);
          double X, yth, yl, mu, q;
          espace_xy_node<PATH>(E[u], nb[j], rc, &X, &yth, &yl);
          if (!near_reflection(X)) continue;
          const double p = ah_hot<0>(literal_x(__ldg(m + k), j, rc), yth, yl, &mu, &q);
          acc = fma(keep_term(p, q) ? p : 0.0, __ldg(dm + k) * rc.span, acc);
        }
      }
      E[0] *= e_ratio;
      E[1] *= e_ratio;
    }
  }
  return acc;
}

template <int MODE, int PATH>
__device__ __forceinline__ double tile_sum_fast_e(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                                  const double* __restrict__ dm, const double* __restrict__ etab,
                                                  double e_ratio, double e_weight, int i0, int i1, int n_points) {
  double acc0 = 0.0, acc1 = 0.0;                                     // sum of mu' E_i
  int first_near = 0x7fffffff;
  const int il = n_points - 1;
  const int i_end = min(i1, il);                                     // the row's last point is added after the loop
  const BracketE br = bracket_e_constants(rc.span * rc.inv_dalt);
  const double2* e2 = reinterpret_cast<const double2*>(etab);        // i0 is even, the table is 16-byte aligned
  const int ip_start = (i0 >> 1) + rc.lane0;                         // index of this thread's first pair of points
  int ip = ip_start;
  const int ip_end = i_end >> 1;                                     // pairs whose two points both precede i_end
  const char* nb_bytes = reinterpret_cast<const char*>(nodes - rc.jlo);   // absolute-level view of the staged window
  while (ip < ip_end) {
    const double2 seed = __ldg(e2 + ip);
    double E0 = seed.x, E1 = seed.y;
    const int ip_blk = min(ip_end, ip + kReseed * rc.group);
#pragma unroll 2
    for (; ip < ip_blk; ip += rc.group) {
      const Node& n0 = *reinterpret_cast<const Node*>(nb_bytes + bracket_bytes_e(E0, br));
      const Node& n1 = *reinterpret_cast<const Node*>(nb_bytes + bracket_bytes_e(E1, br));
      double X0, X1, yth0, yth1, yl0, yl1, mu0, mu1, q0, q1;
      espace_xy_node<PATH>(E0, n0, rc, &X0, &yth0, &yl0);
      espace_xy_node<PATH>(E1, n1, rc, &X1, &yth1, &yl1);
      const bool near0 = (MODE == 0) && near_reflection(X0), near1 = (MODE == 0) && near_reflection(X1);
      const double p0 = ah_hot<MODE>(X0, yth0, yl0, &mu0, &q0);
      const double p1 = ah_hot<MODE>(X1, yth1, yl1, &mu1, &q1);
      if (MODE == 0) {
        if (near0 || near1) first_near = min(first_near, ip);       // left to near_reflection_tail_e
        if (!near0) add_kept(acc0, p0, q0, E0);                     // nansum (lib:288)
        if (!near1) add_kept(acc1, p1, q1, E1);
      } else {
        add_kept(acc0, p0, q0, E0);
        add_kept(acc1, p1, q1, E1);
      }
      E0 *= e_ratio;
      E1 *= e_ratio;
    }
  }
  double acc_km = 0.0;
  if (MODE == 0 && first_near != 0x7fffffff)
    acc_km = near_reflection_tail_e<PATH>(nodes, rc, m, dm, etab, e_ratio, ip_start, first_near, ip_end);
  // The pair that holds the row's last point (lib:416: it weighs 1e-6 km) and, when n_points is even, the point before
  // it (a regular point the loop above left out with its pair): their owner adds them from the tables.
  const int ipl = il >> 1;
  if (2 * ipl >= i0 && 2 * ipl < i1 && (ipl - (i0 >> 1)) % rc.group == rc.lane0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = 2 * ipl + u;
      if (k > il) continue;
      const double Ek = __ldg(etab + k);
      const int jk = min(max(bracket_floor_e(Ek, br), rc.jlo), rc.jhi);
      double Xk, ythk, ylk, muk, qk;
      espace_xy_node<PATH>(Ek, nodes[jk - rc.jlo], rc, &Xk, &ythk, &ylk);
      // (coarse grids: the last point is the only one this close, and with its weight of 1e-6 km an ulp of X is worth
      //  < 1e-11 of the virtual height -- not worth an IEEE sqrt and two divisions on one lane of a 200-point row)
      if (MODE == 0 && (k < il || n_points >= 1024) && near_reflection(Xk)) Xk = literal_x(__ldg(m + k), jk, rc);
      const double pk = ah_hot<MODE>(Xk, ythk, ylk, &muk, &qk);
      acc_km = fma(keep_term(pk, qk) ? pk : 0.0, (k == il) ? kBackoff : __ldg(dm + k) * rc.span, acc_km);
    }
  }
  return fma(acc0 + acc1, e_weight * rc.span, acc_km);
}

// Grid points [i0, i1) of one row on the evaluation path chosen for its profile; returns this thread's share
// of the nansum.  Shared by the tile kernel (group = CTA, nodes pre-scaled for the row) and the row-per-warp
// kernel (group = warp, ROWSCALE).
// Out-of-line copies of the rarely used evaluation paths (non-finite inputs, literal flag, unmagnetised
// profiles, rows clamped to level 0): keeps the hot fast-path loop compact in the instruction cache.
template <int MODE, int PATH>
__device__ __noinline__ double tile_sum_cold(const Node* nodes, const RowConst& rc, const double* __restrict__ m,
                                             int i0, int i1, int np) {
  return tile_sum<MODE, PATH, false>(nodes, rc, m, i0, i1, np, 0.0);
}
template <int MODE, bool LITERAL>
__device__ __noinline__ double const_mup_sum(const Node* nodes, const RowConst& rc, int path, double den0, double b0,
                                             double psi0, const double* __restrict__ m, int i0, int i1, int np) {
  // one evaluation at level 0 with the sign-safe / literal arithmetic, then sum(mu' * dh_i)
  const double X = x_literal(den0, rc.f_hz);
  double mup0;
  if (path == kPathIso) {
    mup0 = iso_mup(X, nullptr);
  } else {
    const double Y = y_literal(b0, rc.f_hz);
    if (LITERAL) {
      mup0 = ah_literal<MODE>(X, Y, psi0, nullptr);
    } else {
      double sn, cs;
      sincos(__dmul_rn(psi0, kDeg2Rad), &sn, &cs);
      mup0 = ah_fast<MODE>(X, Y, sn, cs, nullptr);
    }
  }
  if (mup0 != mup0) return 0.0;                           // every term NaN: nansum drops them all (lib:288)
  if (np >= 2 && isfinite(mup0)) {
    // mu' is one number, so the row's sum is mu' * sum(dh_i), and the weights telescope: dh_i = h_{i+1} - h_i
    // (lib:415) over [i0, i1) adds up to h(i1) - h(i0), plus the final 1e-6 (lib:416) in the row's last tile.  On the
    // rows this path exists for -- reflection at/below the first level, h_c = alt0 - 1e-6 -- all h_i lie within 1e-6 km
    // of each other, every difference is exact (Sterbenz) and so is their sum; what is left, mu' * (h_c - alt0 + 1e-6)
    // ~ mu' * 4e-15 km, is the residue SURVEY.md 7/2 describes (the reference returns alt_min to the last bit or two).
    // 20 % of the reflecting rows of the global grid are of this kind; they used to walk all n_points grid points.
    // (An infinite mu' -- mu == 0 exactly -- keeps the term-by-term loop: its +inf and -inf terms make NaN.)
    if (rc.lane0 != 0) return 0.0;
    const bool last = (i1 >= np);
    const double h_a = __dadd_rn(__dmul_rn(__ldg(m + i0), rc.span), rc.alt0);              // lib:413
    const double h_b = __dadd_rn(__dmul_rn(__ldg(m + (last ? np - 1 : i1)), rc.span), rc.alt0);
    const double term = mup0 * __dsub_rn(h_b, h_a);
    return last ? fma(mup0, kBackoff, term) : term;
  }
  return tile_sum<MODE, kPathFast0, true>(nodes, rc, m, i0, i1, np, mup0);
}

// Grid points [i0, i1) of one row on the evaluation path chosen for its profile; returns this thread's share
// of the nansum.  Shared by the tile kernel (group = CTA, nodes pre-scaled for the row) and the row-per-warp
// kernel (group = warp, ROWSCALE).
struct StretchTables {                                      // p.mult, p.dmult, p.etab and the two ratios
  const double *m, *dm, *e;
  double e_ratio, e_weight;
};
template <int MODE, bool LITERAL, bool ROWSCALE>
__device__ __forceinline__ double row_points(const Node* nodes, const RowConst& rc, int flags, int path, bool const_mup,
                                             double den0, double b0, double psi0, const StretchTables& st, int space,
                                             int i0, int i1, int np) {
  const double* __restrict__ m = st.m;
  const double* __restrict__ dm = st.dm;
  if (const_mup) return const_mup_sum<MODE, LITERAL>(nodes, rc, path, den0, b0, psi0, m, i0, i1, np);
  if (path < kPathGeneral) {
    if (!ROWSCALE) {                                        // nodes staged for this row alone
      if (space == kSpaceE) {                               // uniform altitude grid: E-space loop, no table reads
        if (path == kPathFast0)
          return tile_sum_fast_e<MODE, kPathFast0>(nodes, rc, m, dm, st.e, st.e_ratio, st.e_weight, i0, i1, np);
        if (path == kPathFastS)
          return tile_sum_fast_e<MODE, kPathFastS>(nodes, rc, m, dm, st.e, st.e_ratio, st.e_weight, i0, i1, np);
        return tile_sum_fast_e<MODE, kPathFastL>(nodes, rc, m, dm, st.e, st.e_ratio, st.e_weight, i0, i1, np);
      }
      // m-space loop; uniform grids with more than one staged level take the branch-free bracket
      const bool uni = (flags & kFlagUniformAlt) != 0 && rc.jhi > rc.jlo;
      if (path == kPathFast0)
        return uni ? tile_sum_fast_m<MODE, kPathFast0, true>(nodes, rc, m, dm, i0, i1, np)
                   : tile_sum_fast_m<MODE, kPathFast0, false>(nodes, rc, m, dm, i0, i1, np);
      if (path == kPathFastS)
        return uni ? tile_sum_fast_m<MODE, kPathFastS, true>(nodes, rc, m, dm, i0, i1, np)
                   : tile_sum_fast_m<MODE, kPathFastS, false>(nodes, rc, m, dm, i0, i1, np);
      return uni ? tile_sum_fast_m<MODE, kPathFastL, true>(nodes, rc, m, dm, i0, i1, np)
                 : tile_sum_fast_m<MODE, kPathFastL, false>(nodes, rc, m, dm, i0, i1, np);
    }
    const bool uni = (flags & kFlagUniformAlt) != 0 && rc.jhi > rc.jlo;
    if (path == kPathFast0)
      return uni ? tile_sum_fast<MODE, kPathFast0, true, ROWSCALE>(nodes, rc, m, i0, i1, np)
                 : tile_sum_fast<MODE, kPathFast0, false, ROWSCALE>(nodes, rc, m, i0, i1, np);
    if (path == kPathFastS)
      return uni ? tile_sum_fast<MODE, kPathFastS, true, ROWSCALE>(nodes, rc, m, i0, i1, np)
                 : tile_sum_fast<MODE, kPathFastS, false, ROWSCALE>(nodes, rc, m, i0, i1, np);
    return uni ? tile_sum_fast<MODE, kPathFastL, true, ROWSCALE>(nodes, rc, m, i0, i1, np)
               : tile_sum_fast<MODE, kPathFastL, false, ROWSCALE>(nodes, rc, m, i0, i1, np);
  }
  if (path == kPathIso) return tile_sum_cold<MODE, kPathIso>(nodes, rc, m, i0, i1, np);
  if (LITERAL) return tile_sum_cold<MODE, kPathLiteral>(nodes, rc, m, i0, i1, np);
  return tile_sum_cold<MODE, kPathGeneral>(nodes, rc, m, i0, i1, np);
}

// Stage profile levels [k0, k0 + n) into shared memory.  Fast paths: slopes through one fast reciprocal, density
// and field multiplied by (kx, ky) (the row's cp^2/f^2 and g_p/f, or 1 when the nodes are shared between rows);
// other paths: raw values and numpy's slopes.
// m_space (fast paths of the tile kernels): the node's coordinate becomes m_j = (alt_j - alt0) / span and the slopes
// are taken per unit of m (tile_sum_fast_m); span = 0 keeps altitude space.
// `space` (fast paths of the tile kernels, whose nodes are staged for ONE row): kSpaceM -- the node's coordinate becomes
// m_j = (alt_j - alt0) / span and the slopes are taken per unit of m (tile_sum_fast_m); kSpaceE -- the coordinate becomes
// E_j = (A - m_j) / B, the exponential of the stretched grid at the level, slopes per unit of E (tile_sum_fast_e);
// kSpaceAlt keeps altitude.
__device__ __forceinline__ void stage_nodes(Node* nodes, int k0, int n, int nt, int path, const ProfileRecord& rec,
                                            const double* g_alt, const double* g_den, const double* g_b,
                                            const double* g_psi, double kx, double ky, int tid0, int nthr,
                                            double span = 0.0, int space = kSpaceAlt) {
  const bool fast = path < kPathGeneral;
  if (!fast) space = kSpaceAlt;
  const double inv_span = (space != kSpaceAlt) ? rcp_fast(span) : 0.0;
  for (int q = tid0; q < n; q += nthr) {
    const int k = k0 + q;
    const bool inner = k + 1 < nt;
    // E space: the level at the peak-side end of the profile continues the last segment's line instead of clamping
    // (the unclamped bracket of tile_sum_fast_e may land on it for a point within ~1e-6 levels below it)
    const int kb = (space == kSpaceE && !inner && k > 0) ? k - 1 : k;
    const bool seg = inner || kb != k;
    const double a0 = g_alt[kb], d0 = g_den[kb], b0 = g_b[kb], p0 = g_psi[kb];
    const double a1 = seg ? g_alt[kb + 1] : a0, d1 = seg ? g_den[kb + 1] : d0;
    const double b1 = seg ? g_b[kb + 1] : b0, p1 = seg ? g_psi[kb + 1] : p0;
    Node nd;
    nd.alt = a0;
    if (fast) {
      // slopes through one fast reciprocal (<= 2 ulp from numpy's quotient; the literal paths divide)
      const double inv_dx = seg ? rcp_fast(a1 - a0) : 0.0;
      const bool ext = kb != k;                           // values of level k, slopes of the segment below it
      nd.alt = ext ? a1 : a0;
      nd.x = (ext ? d1 : d0) * kx;
      nd.sx = ((d1 - d0) * kx) * inv_dx;
      nd.y = (ext ? b1 : b0) * ky;
      nd.sy = ((b1 - b0) * ky) * inv_dx;
      nd.srad = ((p1 - p0) * kDeg2Rad) * inv_dx;
      if (space != kSpaceAlt) {
        const double mj = (nd.alt - rec.alt0) * inv_span;
        const double ds = (space == kSpaceE) ? -(span * kStretchB) : span;     // d alt / d coordinate
        nd.alt = (space == kSpaceE) ? (kStretchA - mj) * kStretchDen : mj;
        nd.sx *= ds;
        nd.sy *= ds;
        nd.srad *= ds;
      }
      if (path == kPathFast0) {
        // constant field angle: interpolate YTh = Y sin(psi)/sqrt(2) and YL = Y cos(psi) directly
        const double sh = rec.sn0 * 0.70710678118654752, cc = rec.cs0;
        const double y0 = nd.y, sy0 = nd.sy;
        nd.y = y0 * sh;
        nd.sy = sy0 * sh;
        nd.srad = y0 * cc;
        nd.sn = sy0 * cc;
        nd.cs = 0.0;
        if (space == kSpaceE) {
          // E-space, constant field angle: intercept form.  The three interpolants become ONE FMA each on E itself,
          // v(E) = v_j + s (E - E_j) = (v_j - s E_j) + s E, which drops the subtraction t = E - E_j from every grid point.
          // Safe here and only here: the intercept differs from v_j by |s E_j|, and in E-space that product is small
          // exactly where the integrand is sensitive -- E_j <= ~100 in the segment below the reflection level, which
          // holds half of the row's points (B |dX/dm| E_j ~ 5e-3 against X ~ 1) -- while at the bottom of the row, where
          // it reaches |dX/dm| ~ 1, mu' ~ 1 is insensitive to an X off by 1e-16.  Layout (three aligned LDS.128):
          // {alt, x} = {X intercept, X slope}, {sx, y} = {YTh intercept, slope}, {sy, srad} = {YL intercept, slope}.
          const double ej = nd.alt;
          Node q;
          q.alt = fma(-nd.sx, ej, nd.x);
          q.x = nd.sx;
          q.sx = fma(-nd.sy, ej, nd.y);
          q.y = nd.sy;
          q.sy = fma(-nd.sn, ej, nd.srad);
          q.srad = nd.sn;
          q.sn = ej;                                      // (kept for inspection; the loop does not read it)
          q.cs = 0.0;
          nd = q;
        }
      } else {
        const bool quadratic = (space == kSpaceE && path == kPathFastS);
        if (!quadratic) sincos((ext ? p1 : p0) * kDeg2Rad, &nd.sn, &nd.cs);
        if (quadratic) {
          // E-space, field angle turning by <= 4e-4 rad per level (every IGRF profile): the two field terms
          //   YTh = Y sin(psi) / sqrt(2),  YL = Y cos(psi),   Y and psi linear in E within the segment
          // as QUADRATICS in E, intercept form: two FMAs each instead of the interpolation of Y, the second-order
          // rotation and the two products (13 -> 5 FP64 instructions per point for the three interpolants).
          // The Taylor expansion is centred at the UPPER end of the segment -- the next level, or the row's own top
          // point E = 1 in the segment that holds the reflection level -- because that is where the integrand is
          // singular: a relative error eps of the field terms at the top of the row moves the virtual height by
          // ~3000 eps (an expansion about the segment's lower level, 3.5e-12 off at the far end on the tutorial's Night
          // profile, cost 1e-8).  Dropped: the cubic terms, (dB/B) (d psi)^2 / 2 + (d psi)^3 / 6 relative at the segment's
          // LOWER end (<= 1e-11 + 1.1e-11 by the limits the row setup enforces for this path), falling off with the cube
          // of the distance from the centre.  Cancellation of the intercept form: as for the constant-angle path.
          const double ej = nd.alt;
          const double e_up = seg && !ext ? (kStretchA - (a1 - rec.alt0) * inv_span) * kStretchDen : ej;
          const double ec = fmax(e_up, 1.0);
          const double tc = ec - ej;
          const double y = fma(nd.sy, tc, nd.y), sy = nd.sy, r = nd.srad;
          double sk, ck;
          sincos(fma(r, tc, (ext ? p1 : p0) * kDeg2Rad), &sk, &ck);
          const double h = 0.70710678118654752;
          const double a0 = (y * sk) * h, a1c = (fma(sy, sk, (y * ck) * r)) * h;
          const double a2 = (fma(sy * ck, r, -0.5 * ((y * sk) * (r * r)))) * h;
          const double b0 = y * ck, b1c = fma(sy, ck, -((y * sk) * r));
          const double b2 = -fma(sy * sk, r, 0.5 * ((y * ck) * (r * r)));
          Node q;
          q.alt = fma(-nd.sx, ej, nd.x);                  // X: intercept, slope
          q.x = nd.sx;
          q.sx = fma(fma(a2, ec, -a1c), ec, a0);          // YTh(E) = c0 + c1 E + c2 E^2
          q.y = fma(-2.0 * a2, ec, a1c);
          q.sy = a2;
          q.srad = fma(fma(b2, ec, -b1c), ec, b0);        // YL(E)
          q.sn = fma(-2.0 * b2, ec, b1c);
          q.cs = b2;
          nd = q;
        }
      }
    } else {
      double sd = 0.0, sb = 0.0, sp = 0.0;
      if (inner) {                                        // numpy: slopes[k] = (fp[k+1]-fp[k])/(xp[k+1]-xp[k])
        const double dx = __dsub_rn(a1, a0);
        sd = __ddiv_rn(__dsub_rn(d1, d0), dx);
        sb = __ddiv_rn(__dsub_rn(b1, b0), dx);
        sp = __ddiv_rn(__dsub_rn(p1, p0), dx);
      }
      nd.x = d0; nd.sx = sd; nd.y = b0; nd.sy = sb; nd.srad = sp; nd.sn = p0; nd.cs = 0.0;
    }
    nodes[q] = nd;
  }
}

// ProfileRecord through L2 (ld.cg): written by the row-setup grid, which may still be draining under PDL.
__device__ __forceinline__ ProfileRecord load_profile_record(const ProfileRecord* ptr) {
  union { ProfileRecord r; int4 v[4]; } u;
  const int4* src = reinterpret_cast<const int4*>(ptr);
#pragma unroll
  for (int k = 0; k < 4; ++k) u.v[k] = __ldcg(src + k);
  return u.r;
}

// Evaluation path of a profile from the flags the row setup recorded.
template <bool LITERAL>
__device__ __forceinline__ int select_path(int flags) {
  if (flags & kFlagIso) return kPathIso;
  if (LITERAL) return kPathLiteral;
  if (flags & kFlagGeneral) return kPathGeneral;
  if (flags & kFlagPsiConst) return kPathFast0;
  if (flags & kFlagPsiSmall) return kPathFastS;
  return kPathFastL;
}

// Block reduction of a tile's share, combination of the segments of a row, lib:288-292.
__device__ __forceinline__ void finish_tile(const VfoParams& p, BlockScratch& sc, double acc, int64_t lrow, int seg,
                                            int n_seg, int64_t out_idx, double alt_min) {
  PRHF_TRACE_MARK(6);
  const double s_tile = block_sum(acc, sc);
  PRHF_TRACE_MARK(7);
#ifdef PRHF_TRACE
  if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 8 + 0] |= (trace_globaltimer() << 10);
#endif
  if (threadIdx.x != 0) return;
  double total = s_tile;
  if (n_seg > 1) {
    double* part = p.partial + lrow * p.max_seg;
    __stcg(part + seg, s_tile);
    __threadfence();
    const unsigned prev = atomicAdd(p.counter + lrow, 1u);
    if (prev != (unsigned)(n_seg - 1)) return;
    __threadfence();
    total = 0.0;
    for (int s = 0; s < n_seg; ++s) total += __ldcg(part + s);     // fixed order: deterministic
    p.counter[lrow] = 0u;                                 // self-reset for the next launch
  }
  if (total == 0.0) total = CUDART_NAN;                   // lib:290
  p.vh[out_idx] = total + alt_min;                        // lib:292
}

// The single-launch kernel has the profile's raw levels in shared memory already (its row setup staged them in the
// first 4 * n_alt doubles of the dynamic shared memory): the tile then builds its nodes from those instead of loading
// the levels from global memory a second time, and puts the nodes behind them -- when the tile's window fits the
// `upper_bytes` that are left (it does whenever the window is at most half the altitude grid; else global loads).
struct TileSrc {
  const double *alt, *den, *b, *psi;
  unsigned upper_off, upper_bytes;
};

// Queued mode: the end of a whole-row tile in ONE barrier.  The warps leave their partial sums in `part` (the caller
// alternates two buffers from tile to tile), thread 0 publishes the next tile's index alongside, and after the barrier
// thread 0 alone adds the partials in warp order (deterministic) and writes the row while the other warps are already
// in the next tile's prologue.  The general finish_tile costs two barriers plus the caller's end-of-tile barrier.
struct QueueFinish {
  double* part;            // [kMaxWarps]
  unsigned* next_slot;     // shared word the tile loop reads its next index from
  unsigned next_value;     // valid in thread 0
  int cap_nodes;           // levels the CTA's node buffer holds; a row whose window is larger is deferred:
  unsigned* defer_count;   // ... appended to this list for the full-width kernel that follows
  LiveRow* defer_list;
};

// One tile: grid points [seg * seg_len, (seg+1) * seg_len) of row `lrow`.
template <int MODE, bool LITERAL, int NT = kTileThreads>
__device__ __forceinline__ void tile_body(const VfoParams& p, const int64_t lrow, const double span,
                                          const ProfileRecord* rec_src, const int seg, const int n_seg,
                                          const int seg_len, unsigned char* smem_raw, BlockScratch& sc,
                                          const TileSrc* src = nullptr, const QueueFinish* qf = nullptr) {
  const int tid = threadIdx.x;
#ifdef PRHF_TRACE
  if (p.trace && tid == 0) {
    p.trace[(size_t)blockIdx.x * 8 + 0] = trace_smid();
    p.trace[(size_t)blockIdx.x * 8 + 1] = trace_globaltimer();
  }
#endif
  PRHF_TRACE_MARK(2);
  // a launch covers < 2^24 rows (host side), so 32-bit division is enough (the 64-bit form costs ~3 % of the
  // kernel's instructions: profiles/ncu_r01h_tile_kernel_batch256_lines.txt)
  const unsigned lprof32 = (unsigned)lrow / (unsigned)p.n_freq;
  const int r = (int)((unsigned)lrow - lprof32 * (unsigned)p.n_freq);
  const int64_t lprof = lprof32;
  const int64_t prof = p.profile_offset + lprof;
  const int i0 = seg * seg_len;
  const int i1 = min(p.n_points, i0 + seg_len);
  // every load of the prologue is independent of the others: issue them together
  const ProfileRecord rec = rec_src ? *rec_src : load_profile_record(p.prof_rec + lprof);
  const double f_mhz = p.freq[prof * p.freq_stride + r];
  // (a tile that is the whole row starts at m = 0, E = e^10 and ends at m = 1, E = 1: no table reads on its prologue)
  const bool whole_row = (n_seg == 1) && p.n_points > 1;
  const double m_lo = whole_row ? 0.0 : __ldg(p.mult + i0), m_hi = whole_row ? 1.0 : __ldg(p.mult + i1 - 1);
  const double e_lo = whole_row ? 22026.465794806718 : __ldg(p.etab + i0);
  const double e_hi = whole_row ? 1.0 : __ldg(p.etab + i1 - 1);
  if (!(span == span)) return;                            // no reflection / failed profile: K1 wrote the NaN
  PRHF_TRACE_MARK(3);
  const int nt = rec.nt;

  const int path = select_path<LITERAL>(rec.flags);
  const int A = p.n_alt;
  const double* g_den = p.den + prof * A;
  const double* g_b = p.bmag + prof * A;
  const double* g_psi = p.bpsi + prof * A;
  const double* g_alt = p.alt + prof * p.alt_stride;

  RowConst rc;
  rc.g_alt = g_alt;
  rc.g_den = g_den;
  rc.f_hz = __dmul_rn(f_mhz, 1e6);
  rc.alt0 = rec.alt0;
  rc.span = span;
  rc.inv_dalt = rec.inv_dalt;
  rc.nt = nt;
  double kx, ky;                                          // fast paths: X = den * kx, Y = b * ky
  row_scales(rc.f_hz, &kx, &ky);
  const bool unsorted = (rec.flags & kFlagAltUnsorted) != 0;
  rc.unsorted = unsorted ? 1 : 0;
  const bool const_mup = (!(span > 0.0) || nt == 1) && !unsorted;   // h_c <= alt0: every point clamps to level 0
  PRHF_TRACE_X(10);

  // ---- node window of this tile: brackets of its first and last grid point ----
  int space = kSpaceM;                                    // coordinate the fast paths stage their nodes in
  if (const_mup) {
    rc.jlo = rc.jhi = 0;
  } else if (unsorted) {
    rc.jlo = 0;                                           // np.interp's range tests need both ends of the axis
    rc.jhi = nt - 1;
  } else {
    const double h_lo = __dadd_rn(__dmul_rn(m_lo, span), rc.alt0);
    const double h_hi = __dadd_rn(__dmul_rn(m_hi, span), rc.alt0);
    const int g_lo = min(max(__double2int_rd((h_lo - rc.alt0) * rec.inv_dalt), 0), nt - 1);
    const int g_hi = min(max(__double2int_rd((h_hi - rc.alt0) * rec.inv_dalt), 0), nt - 1);
    if (rec.flags & kFlagUniformAlt) {
      // levels sit within a quarter step of the uniform grid: the guess is the bracket to +-1.  The m-space loop takes
      // floor(m_i * c1) unclamped (bracket_floor_m), so the window must contain that value for the tile's first and last
      // multiplier -- computed here with the same instruction -- and, by monotonicity, for every point in between.
      const double c1 = span * rec.inv_dalt;
      const int f_lo = bracket_floor_m(m_lo, c1), f_hi = bracket_floor_m(m_hi, c1);
      rc.jlo = max(min(g_lo - 1, f_lo), 0);
      rc.jhi = min(max(g_hi + 1, f_hi), nt - 1);
#ifndef PRHF_NO_ESPACE                                    // (developer A/B switch: keep the table-reading m-space loop)
      if (path < kPathGeneral && nt <= kBracketMaxLevels && rc.jhi > rc.jlo) {
        // E-space loop (tile_sum_fast_e): its bracket, evaluated at the tile's first and last point, +-1 level for the
        // drift of the recurrence against the table (1e-14 relative)
        space = kSpaceE;
        const BracketE br = bracket_e_constants(c1);
        rc.jlo = max(min(rc.jlo, bracket_floor_e(e_lo, br) - 1), 0);
        rc.jhi = min(max(rc.jhi, bracket_floor_e(e_hi, br) + 1), nt - 1);
      }
#endif
    } else {
      if (tid == 0 || tid == 32) {                        // NT >= 64
        const double h = (tid == 0) ? h_lo : h_hi;
        const int gj = (tid == 0) ? g_lo : g_hi;
        int j;
        if (h >= g_alt[gj] && (gj == nt - 1 || h < g_alt[gj + 1])) j = gj;
        else j = max(bracket_in<1>(h, g_alt, 0, nt - 1), 0);
        sc.bcast_i[tid >> 5] = j;
      }
      __syncthreads();
      rc.jlo = min(sc.bcast_i[0], sc.bcast_i[1]);
      rc.jhi = max(sc.bcast_i[0], sc.bcast_i[1]);
    }
  }
  PRHF_TRACE_MARK(4);
  const int n_stage = min(rc.jhi + 1, nt - 1) - rc.jlo + 1;          // levels jlo .. min(jhi+1, nt-1)

  Node* nodes = reinterpret_cast<Node*>(smem_raw);
  if (const_mup) space = kSpaceAlt;
  if (qf && qf->cap_nodes > 0 && n_stage > qf->cap_nodes) {
    // narrow queue kernel: the row's window does not fit this CTA's node buffer -- hand the row to the full-width
    // kernel that follows (rows that reflect in the upper half of a long altitude grid; block-uniform decision)
    if (tid == 0) {
      LiveRow e;
      e.row = (int)lrow;
      e.pad = 0;
      e.span = span;
      qf->defer_list[atomicAdd(qf->defer_count, 1u)] = e;
      *qf->next_slot = qf->next_value;
    }
    __syncthreads();
    return;
  }
  if (src && (size_t)n_stage * sizeof(Node) <= src->upper_bytes) {
    // raw levels from shared memory, nodes behind them (single-launch kernel)
    nodes = reinterpret_cast<Node*>(smem_raw + src->upper_off);
    stage_nodes(nodes, rc.jlo, n_stage, nt, path, rec, src->alt, src->den, src->b, src->psi, kx, ky, tid, NT, span, space);
  } else {
    if (src) __syncthreads();                             // (every thread is done with the raw levels in shared memory)
    stage_nodes(nodes, rc.jlo, n_stage, nt, path, rec, g_alt, g_den, g_b, g_psi, kx, ky, tid, NT, span, space);
  }
  __syncthreads();
  PRHF_TRACE_MARK(5);

  // ---- grid points of the tile ----
  rc.lane0 = tid;
  rc.group = NT;
  rc.kx = kx;
  rc.ky = ky;
  const StretchTables st{p.mult, p.dmult, p.etab, p.e_ratio, p.e_weight};
  const double acc = row_points<MODE, LITERAL, false>(nodes, rc, rec.flags, path, const_mup, g_den[0], g_b[0], g_psi[0],
                                                      st, space, i0, i1, p.n_points);

  // ---- reduce, finish (lib:288-292) ----
  if (qf) {                                               // queued mode, n_seg == 1
    const double w = warp_sum(acc);
    if ((tid & 31) == 0) qf->part[tid >> 5] = w;
    if (tid == 0) *qf->next_slot = qf->next_value;
    __syncthreads();
    if (tid == 0) {
      double total = 0.0;
#pragma unroll
      for (int k = 0; k < NT / 32; ++k) total += qf->part[k];
      if (total == 0.0) total = CUDART_NAN;               // lib:290
      p.vh[prof * p.n_freq + r] = total + rec.alt_min;    // lib:292
    }
    return;
  }
  finish_tile(p, sc, acc, lrow, seg, n_seg, prof * p.n_freq + r, rec.alt_min);
}

// Direct mode (large batches): tile = blockIdx.x, n_seg fixed by the host; rows without reflection exit.
// Planned mode (small batches): K1 appended the rows that reflect to a compact list.  With few rows the
// kernel is latency-bound unless only live rows get tiles and the segment count makes the live tiles fill
// the resident-CTA slots, so every CTA sizes the tiling from the live-row count (same arithmetic in every
// CTA, candidates prepared by the host) and strides over live_rows * n_seg tiles.
template <int MODE, bool LITERAL>
__device__ __forceinline__ void planned_tiles(const VfoParams& p, unsigned char* smem_raw, BlockScratch& sc) {
  const int live = (int)__ldcg(p.live_count);
  // How many segments per live row?  Measured (profiles/sweep_nseg_batch_r01.log, 2 ... 23 profiles): the step is
  // shortest when the live tiles number about 1.3 x the resident CTA slots -- enough to even out where the hardware
  // places the CTAs (under programmatic dependent launch they arrive while the row-setup grid still holds part of
  // the SMs) -- as long as a tile keeps >= kPlanMinTilePoints grid points to amortise its prologue.  A cost model
  // with per-SM tile counts was tried first and proved brittle: its ceil() steps flip the choice on a 5 % change of
  // the live-row count.
  int n_seg = 1, seg_len = p.n_points;
  {
    const float want = kPlanTilesPerSlot * (float)p.slots / (float)max(live, 1);
    float best = 3.0e38f;
    for (int c = 0; c < p.n_cand; ++c) {
      if (p.cand_seg[c] > 1 && p.cand_len[c] < kPlanMinTilePoints) continue;
      const float d = fabsf((float)p.cand_seg[c] - want);
      if (d < best) { best = d; n_seg = p.cand_seg[c]; seg_len = p.cand_len[c]; }
    }
  }
  const int n_tiles = live * n_seg;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int li = t / n_seg;
    // written earlier in this launch sequence (possibly in this very kernel): bypass L1
    const int4 raw = __ldcg(reinterpret_cast<const int4*>(p.live_list + li));
    const double span = __hiloint2double(raw.w, raw.z);
    tile_body<MODE, LITERAL>(p, raw.x, span, nullptr, t - li * n_seg, n_seg, seg_len, smem_raw, sc);
    __syncthreads();                                      // shared memory is reused by the next tile
  }
}

template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks) vfo_tile_kernel(const VfoParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BlockScratch sc;
  // While the row-setup grid is still running (PDL): pull the multiplier table into L2.  It does not depend
  // on K1, and after an L2 flush its first touch would otherwise be a DRAM miss inside the grid loop.
  if (p.live_count != nullptr) {
    const size_t bytes = sizeof(double) * 2 * mult_table_len(p.n_points);      // [m | dm], contiguous
    for (size_t off = ((size_t)blockIdx.x * kTileThreads + threadIdx.x) * 128; off < bytes;
         off += (size_t)gridDim.x * kTileThreads * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.mult) + off));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the row-setup grid has completed (no-op without PDL)
  if (p.live_count == nullptr) {
    const unsigned tile = blockIdx.x;                     // grid.x itself is 32-bit
    const unsigned lrow = (p.n_seg == 1) ? tile : tile / (unsigned)p.n_seg;
    // Two rows out of three of a global batch do not reflect: their CTAs leave on the row's span before anything
    // else is computed or loaded (the row-setup kernel already wrote their NaN).
    const double span = p.row_span[lrow];
    if (!(span == span)) return;
    const int seg = (p.n_seg == 1) ? 0 : (int)(tile - lrow * (unsigned)p.n_seg);
    tile_body<MODE, LITERAL>(p, lrow, span, nullptr, seg, p.n_seg, p.seg_len, smem_raw, sc);
    return;
  }
  planned_tiles<MODE, LITERAL>(p, smem_raw, sc);
}

// Queued mode (large batches): one CTA per resident slot draws whole-row tiles from the queue the row setup filled.
// * Tiles are handed out by ticket.  A static stride over the queue measured 11 % SLOWER than one CTA per row
//   (profiles/queue_modes_r02.txt): with equal tile counts the launch ends with its slowest SM, whereas the hardware's
//   block scheduler -- and a ticket -- give a faster SM more tiles.  The ticket for the NEXT tile is drawn before the
//   current one is computed, so its latency hides behind the grid loop.
// * Narrow CTAs: 128 threads, eight CTAs per SM instead of four of 256.  A tile is a whole row either way, so the
//   per-tile prologue and the barriers at its ends cost the same number of cycles but half the share of a CTA's
//   residency, and a CTA in its prologue idles 4 warps of 32 instead of 8 (+3 %, profiles/queue_modes_r02.txt).  Eight
//   CTAs leave each of them p.queue_cap_nodes levels of node buffer (434 on B200); a row whose window is larger is
//   deferred to the full-width kernel, launched behind this one over the deferred list.
#ifndef PRHF_QUEUE_THREADS
#define PRHF_QUEUE_THREADS 128
#endif
#ifndef PRHF_QUEUE_MINB
#define PRHF_QUEUE_MINB 8
#endif
constexpr int kQueueThreads = PRHF_QUEUE_THREADS;         // (developer builds: make ab NAME=.. DEFS="-DPRHF_QUEUE_MINB=10")
constexpr int kQueueMinBlocks = PRHF_QUEUE_MINB;
template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kQueueThreads, kQueueMinBlocks) vfo_queue_kernel(const VfoParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BlockScratch sc;
  __shared__ unsigned s_next;
  __shared__ double s_part[2][kMaxWarps];
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the row-setup grid has completed (no-op without PDL)
  const unsigned n_tiles = __ldcg(p.live_count);
  unsigned t = blockIdx.x, flip = 0;
  while (t < n_tiles) {
    unsigned ticket = 0;
    if (threadIdx.x == 0) ticket = atomicAdd(p.live_count + 1, 1u);
    const int4 raw = __ldcg(reinterpret_cast<const int4*>(p.live_list + t));
    const double span = __hiloint2double(raw.w, raw.z);
    QueueFinish qf;
    qf.part = s_part[flip];
    qf.next_slot = &s_next;
    qf.next_value = ticket + gridDim.x;
    qf.cap_nodes = p.queue_cap_nodes;
    qf.defer_count = p.live_count + 2;
    qf.defer_list = p.defer_list;
    flip ^= 1u;
    tile_body<MODE, LITERAL, kQueueThreads>(p, raw.x, span, nullptr, 0, 1, p.n_points, smem_raw, sc, nullptr, &qf);
    t = s_next;
  }
}

// Row-per-warp form for small n_points (direct mode): one CTA stages the profile's levels ONCE (un-scaled)
// and its warps each take whole rows, so neither the staging (~100 FP64 instructions per level) nor a
// block-wide barrier is paid per 200-point row.  Rows that do not reflect are skipped on their row_span.
template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks) vfo_rowwarp_kernel(const VfoParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the row-setup grid has completed (no-op without PDL)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunks = (p.n_freq + p.rw_rows_per_cta - 1) / p.rw_rows_per_cta;
  const int64_t lprof = blockIdx.x / chunks;
  const int r_begin = (int)(blockIdx.x % chunks) * p.rw_rows_per_cta;
  const int r_end = min(p.n_freq, r_begin + p.rw_rows_per_cta);
  const int64_t prof = p.profile_offset + lprof;
  const ProfileRecord rec = load_profile_record(p.prof_rec + lprof);
  if (rec.flags & kFlagFailed) return;                    // K1 wrote the NaN rows and the status
  // anything to do in this chunk?
  bool live = false;
  for (int r = r_begin + tid; r < r_end; r += kTileThreads) {
    const double sp = p.row_span[lprof * p.n_freq + r];
    live |= (sp == sp);
  }
  if (!__syncthreads_or(live)) return;

  const int nt = rec.nt;
  const int path = select_path<LITERAL>(rec.flags);
  const bool unsorted = (rec.flags & kFlagAltUnsorted) != 0;

  const int A = p.n_alt;
  const double* g_den = p.den + prof * A;
  const double* g_b = p.bmag + prof * A;
  const double* g_psi = p.bpsi + prof * A;
  const double* g_alt = p.alt + prof * p.alt_stride;
  Node* nodes = reinterpret_cast<Node*>(smem_raw);
  stage_nodes(nodes, 0, nt, nt, path, rec, g_alt, g_den, g_b, g_psi, 1.0, 1.0, tid, kTileThreads);
  __syncthreads();
  const double den0 = g_den[0], b0 = g_b[0], psi0 = g_psi[0];

  for (int r = r_begin + wid; r < r_end; r += kTileThreads / 32) {
    const int64_t lrow = lprof * p.n_freq + r;
    const double span = p.row_span[lrow];
    if (!(span == span)) continue;
    RowConst rc;
    rc.g_alt = g_alt;
    rc.g_den = g_den;
    rc.f_hz = __dmul_rn(p.freq[prof * p.freq_stride + r], 1e6);
    rc.alt0 = rec.alt0;
    rc.span = span;
    rc.inv_dalt = rec.inv_dalt;
    rc.nt = nt;
    row_scales(rc.f_hz, &rc.kx, &rc.ky);
    rc.lane0 = lane;
    rc.group = 32;
    rc.unsorted = unsorted ? 1 : 0;
    const bool const_mup = (!(span > 0.0) || nt == 1) && !unsorted;   // h_c <= alt0: every point clamps to level 0
    rc.jlo = 0;
    rc.jhi = const_mup ? 0 : nt - 1;
    const StretchTables st{p.mult, p.dmult, p.etab, p.e_ratio, p.e_weight};
    double acc = row_points<MODE, LITERAL, true>(nodes, rc, rec.flags, path, const_mup, den0, b0, psi0, st, kSpaceAlt,
                                                 0, p.n_points, p.n_points);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (acc == 0.0) acc = CUDART_NAN;                   // lib:290
      p.vh[prof * p.n_freq + r] = acc + rec.alt_min;      // lib:292
    }
  }
}

// ---- profiles with more levels than the shared-memory staging holds (n_alt > prhf_max_n_alt()) ----
// The reference has no such limit (np.interp, lib:424-426).  Row setup: the thread-per-frequency mapping reading the
// levels in place (rows_body, p.levels_in_global).  Then one CTA per profile writes the profile's UN-scaled nodes
// (np.interp slopes, sin/cos of the field angle) to a table in global memory, and the tile kernel below reads them
// through L1/L2 with the row's cp^2/f^2 and g_p/f applied per point -- the arithmetic of the row-per-warp kernel with
// a CTA per tile.  Slower per point than the staged form (two more FP64 multiplies, table reads instead of shared
// memory), but every decision and every value is the same.
template <bool LITERAL>
__global__ void __launch_bounds__(kTileThreads) vfo_nodes_global_kernel(const VfoParams p) {
  const int64_t lprof = blockIdx.x;
  const int64_t prof = p.profile_offset + lprof;
  const ProfileRecord rec = p.prof_rec[lprof];
  if (rec.flags & kFlagFailed) return;
  const int A = p.n_alt;
  Node* nodes = reinterpret_cast<Node*>(p.node_table) + lprof * A;
  stage_nodes(nodes, 0, rec.nt, rec.nt, select_path<LITERAL>(rec.flags), rec, p.alt + prof * p.alt_stride,
              p.den + prof * A, p.bmag + prof * A, p.bpsi + prof * A, 1.0, 1.0, threadIdx.x, kTileThreads);
}

template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kTileThreads, kTileMinBlocks) vfo_tile_global_kernel(const VfoParams p) {
  __shared__ BlockScratch sc;
  const unsigned tile = blockIdx.x;
  const unsigned lrow = (p.n_seg == 1) ? tile : tile / (unsigned)p.n_seg;
  const int seg = (p.n_seg == 1) ? 0 : (int)(tile - lrow * (unsigned)p.n_seg);
  const double span = p.row_span[lrow];
  if (!(span == span)) return;                            // no reflection / failed profile: the row setup wrote the NaN
  const unsigned lprof = lrow / (unsigned)p.n_freq;
  const int r = (int)(lrow - lprof * (unsigned)p.n_freq);
  const int64_t prof = p.profile_offset + lprof;
  const ProfileRecord rec = p.prof_rec[lprof];
  const int A = p.n_alt, nt = rec.nt;
  const int path = select_path<LITERAL>(rec.flags);
  const double* g_den = p.den + prof * A;
  RowConst rc;
  rc.g_alt = p.alt + prof * p.alt_stride;
  rc.g_den = g_den;
  rc.f_hz = __dmul_rn(p.freq[prof * p.freq_stride + r], 1e6);
  rc.alt0 = rec.alt0;
  rc.span = span;
  rc.inv_dalt = rec.inv_dalt;
  rc.nt = nt;
  row_scales(rc.f_hz, &rc.kx, &rc.ky);
  rc.lane0 = threadIdx.x;
  rc.group = kTileThreads;
  rc.unsorted = (rec.flags & kFlagAltUnsorted) ? 1 : 0;
  const bool const_mup = (!(span > 0.0) || nt == 1) && !rc.unsorted;
  rc.jlo = 0;
  rc.jhi = const_mup ? 0 : nt - 1;
  const int i0 = seg * p.seg_len, i1 = min(p.n_points, i0 + p.seg_len);
  const Node* nodes = reinterpret_cast<const Node*>(p.node_table) + (size_t)lprof * A;
  const StretchTables st{p.mult, p.dmult, p.etab, p.e_ratio, p.e_weight};
  const double acc = row_points<MODE, LITERAL, true>(nodes, rc, rec.flags, path, const_mup, g_den[0],
                                                     p.bmag[prof * A], p.bpsi[prof * A], st, kSpaceAlt, i0, i1,
                                                     p.n_points);
  finish_tile(p, sc, acc, lrow, seg, p.n_seg, prof * p.n_freq + r, rec.alt_min);
}

// Solo form for a single profile (rows * n_seg <= resident CTAs): ONE launch, no hand-off through global
// memory.  Every CTA owns one tile = (row, segment), runs the row setup for its own row first (the segments of
// a row repeat it, in parallel) and goes straight on to its grid points.  Removes the serial row-setup kernel
// and the inter-kernel gap from the latency-critical single-profile call.
template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kTileThreads, kSoloMinBlocks) vfo_solo_kernel(const VfoParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ BlockScratch sc;
  __shared__ ProfileRecord s_rec;
  __shared__ double s_span;
  const int64_t tile = blockIdx.x;
  const int64_t lrow = tile / p.n_seg;
  {
    // Everything this CTA will read later from cold memory is requested now, so that the DRAM round trips overlap
    // the profile staging instead of following it one by one: the row's frequency, the ends of the tile's
    // multiplier segment (window computation) and the segment itself (first touch in the grid loop).
    const int seg = (int)(tile % p.n_seg);
    const int i0 = seg * p.seg_len, i1 = min(p.n_points, i0 + p.seg_len);
    const int64_t prof = p.profile_offset + lrow / p.n_freq;
    if (threadIdx.x == 0) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.freq + prof * p.freq_stride + lrow % p.n_freq));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.mult + i0));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.mult + max(i1 - 1, i0)));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.etab + max(i1 - 1, i0)));
    }
    // the seeds of the E-space loop: one pair of table entries per thread, the tile's first 4 KB of E
    if (threadIdx.x >= 32 && threadIdx.x < 64)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.etab + i0 + (threadIdx.x - 32) * 16));
    // the rows of a profile share the segment: each CTA requests only its own slice of the lines, so the segment
    // is pulled into L2 once instead of once per row
    // ... and the second line of the kernel parameters (cold constant bank), whose first reader would otherwise be
    // the row scan
    asm volatile("" ::"d"(p.freq_scale), "l"(p.row_span), "l"(p.partial));
    const char* mseg = reinterpret_cast<const char*>(p.mult + i0);
    const char* dseg = reinterpret_cast<const char*>(p.dmult + i0);
    const int n_lines = (int)(((size_t)(i1 - i0 + 4) * sizeof(double) + 127) / 128);
    const int line = (int)(lrow % p.n_freq) + (int)threadIdx.x * p.n_freq;
    if (line < n_lines) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(mseg + (size_t)line * 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(dseg + (size_t)line * 128));
    }
  }
  if (threadIdx.x == 0) s_span = CUDART_NAN;
  rows_body(p, MODE, lrow, reinterpret_cast<double*>(smem_raw), sc, &s_rec, &s_span);   // syncs internally
  __syncthreads();
  const double span = s_span;
  if (!(span == span)) return;                            // no reflection / failed profile: NaN already written
  const size_t raw_bytes = sizeof(double) * 4 * (size_t)p.n_alt;
  const size_t all_bytes = sizeof(Node) * (size_t)p.n_alt;             // the launcher grants max(5 n_alt doubles, this)
  const double* raw = reinterpret_cast<const double*>(smem_raw);       // rows_body: den | alt | b | psi
  TileSrc src;
  src.den = raw;
  src.alt = raw + p.n_alt;
  src.b = raw + 2 * (size_t)p.n_alt;
  src.psi = raw + 3 * (size_t)p.n_alt;
  src.upper_off = (unsigned)raw_bytes;
  src.upper_bytes = (unsigned)(all_bytes > raw_bytes ? all_bytes - raw_bytes : 0);
  tile_body<MODE, LITERAL>(p, lrow, span, &s_rec, (int)(tile % p.n_seg), p.n_seg, p.seg_len, smem_raw, sc, &src);
}

// ------------------------------------------------------------------------------------------
// elementwise mu / mu' (lib:161-256) for callers outside the fused operator
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
// residual of the inversion objective (residual_VH, lib:660-668): NaN model heights are replaced by
// max(nanmean|vh_model|, 100) (lib:664-665), residual = vh_obs - vh_model (lib:668); chi2 = sum residual^2 is
// what lmfit's brute-force search minimises (lib:794-798).  One warp per candidate profile.
// ------------------------------------------------------------------------------------------
__global__ void residual_kernel(const double* __restrict__ vh, const double* __restrict__ vh_obs, int64_t n_profiles,
                                int n_freq, double* __restrict__ residual, double* __restrict__ chi2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pr = warp; pr < n_profiles; pr += n_warps) {
    const double* row = vh + pr * n_freq;
    double s = 0.0;
    int cnt = 0;
    for (int k = lane; k < n_freq; k += 32) {
      const double v = row[k];
      if (v == v) { s += fabs(v); ++cnt; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    // np.nanmean of an all-NaN row is NaN and np.maximum(NaN, 100) is NaN: the whole residual is NaN
    const double fill = (cnt > 0) ? fmax(s / (double)cnt, 100.0) : CUDART_NAN;
    double c = 0.0;
    for (int k = lane; k < n_freq; k += 32) {
      const double v = row[k];
      const double r = vh_obs[k] - ((v == v) ? v : fill);
      if (residual) residual[pr * n_freq + k] = r;
      c = fma(r, r, c);
    }
    c = warp_sum(c);
    if (lane == 0 && chi2) chi2[pr] = c;
  }
}

// ------------------------------------------------------------------------------------------
// argmin of the brute-force objective (lmfit's brute search at lib:794-798 takes the grid node with the smallest
// sum of squared residuals; scipy.optimize.brute: argmin of the raveled grid, first minimum wins).  NaN scores
// (a candidate whose every model height is NaN, or a failed profile) are skipped; out = {index or -1, value}.
// One CTA: the candidate grids of this path have 10^2 .. 10^5 nodes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) argmin_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  double best = CUDART_INF;
  long long bi = -1;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = v[i];
    if (x == x && (bi < 0 || x < best)) { best = x; bi = i; }     // ascending i per thread: first minimum kept
  }
  __shared__ double s_v[32];
  __shared__ long long s_i[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      const double ov = s_v[k];
      const long long oi = s_i[k];
      if (oi >= 0 && (bi < 0 || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    out[0] = (double)bi;
    out[1] = (bi >= 0) ? best : CUDART_NAN;
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

// accuracy self-test of the fast reciprocal / reciprocal square root (max relative error vs IEEE):
// err[0] rcp_fast, err[1] rsqrt_fast, err[2] raw rcp seed, err[3] raw rsqrt seed,
// err[4] rcp seed + one cubic step, err[5] rsqrt seed + one cubic step
__global__ void math_selftest_kernel(int n, double* __restrict__ err) {
  double e[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // log-uniform samples in [1e-30, 1e30] interleaved with a dense sweep of one binade
    const double u = (double)i / (double)n;
    const double x = (i & 1) ? exp(138.0 * (u - 0.5)) : 1.0 + u;
    const double rc = __drcp_rn(x);
    const double rs = __drcp_rn(__dsqrt_rn(x));
    double y0, z0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z0) : "d"(x));
    const double ey = fma(-x, y0, 1.0);
    const double y1 = fma(y0, fma(ey, ey, ey), y0);
    const double ez = fma(-(x * z0), z0, 1.0);
    const double z1 = fma(z0, ez * fma(ez, 0.375, 0.5), z0);
    e[0] = fmax(e[0], fabs(rcp_fast(x) - rc) * x);
    e[1] = fmax(e[1], fabs(rsqrt_fast(x) - rs) / rs);
    e[2] = fmax(e[2], fabs(y0 - rc) * x);
    e[3] = fmax(e[3], fabs(z0 - rs) / rs);
    e[4] = fmax(e[4], fabs(y1 - rc) * x);
    e[5] = fmax(e[5], fabs(z1 - rs) / rs);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = warp_max(e[k]);
    if ((threadIdx.x & 31) == 0)
      atomicMax(reinterpret_cast<unsigned long long*>(err + k), (unsigned long long)__double_as_longlong(v));
  }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
size_t vfo_rows_smem_bytes(int n_alt) { return sizeof(double) * (4 + kRowsPerCta) * (size_t)n_alt; }
size_t vfo_tile_smem_bytes(int n_alt) { return sizeof(Node) * (size_t)n_alt; }
size_t vfo_smem_bytes(int n_alt) {
  const size_t a = vfo_rows_smem_bytes(n_alt), b = vfo_tile_smem_bytes(n_alt);
  return a > b ? a : b;
}

// cudaFuncSetAttribute costs ~1-2 us of host time per call; remember the largest size already granted per
// (device, kernel) so that steady-state launches skip it (also keeps it out of stream capture).
static cudaError_t grant_dynamic_smem(const void* func, int slot, size_t smem) {
  static size_t granted[64][20] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (smem <= granted[dev][slot]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) granted[dev][slot] = smem;
  return e;
}

size_t vfo_node_bytes() { return sizeof(Node); }

cudaError_t launch_vfo_nodes_global(const VfoParams& p, bool literal, int64_t n_profiles, cudaStream_t stream) {
  if (literal) vfo_nodes_global_kernel<true><<<(unsigned)n_profiles, kTileThreads, 0, stream>>>(p);
  else vfo_nodes_global_kernel<false><<<(unsigned)n_profiles, kTileThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_vfo_tiles_global(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream) {
  const unsigned g = (unsigned)n_tiles;
  if (mode == 0) {
    if (literal) vfo_tile_global_kernel<0, true><<<g, kTileThreads, 0, stream>>>(p);
    else vfo_tile_global_kernel<0, false><<<g, kTileThreads, 0, stream>>>(p);
  } else {
    if (literal) vfo_tile_global_kernel<1, true><<<g, kTileThreads, 0, stream>>>(p);
    else vfo_tile_global_kernel<1, false><<<g, kTileThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_vfo_rows(const VfoParams& p, int mode, int64_t n_profiles, cudaStream_t stream) {
  // (the thread-per-frequency mapping needs no per-warp scratch behind the four staged arrays: more CTAs per SM)
  const size_t smem = p.levels_in_global ? 0
                      : (p.k1_lane_mode ? sizeof(double) * 4 * (size_t)p.n_alt : vfo_rows_smem_bytes(p.n_alt));
  cudaError_t e = grant_dynamic_smem((const void*)vfo_rows_kernel, 0, smem);
  if (e != cudaSuccess) return e;
  const int rows_per_cta = (p.k1_lane_mode || p.levels_in_global) ? kThreads : kRowsPerCta * p.rows_per_warp;
  const int chunks = (p.n_freq + rows_per_cta - 1) / rows_per_cta;
  vfo_rows_kernel<<<(unsigned)(n_profiles * chunks), kThreads, smem, stream>>>(p, mode);
  return cudaGetLastError();
}

template <int MODE, bool LITERAL>
static cudaError_t launch_rowwarp_t(const VfoParams& p, int64_t n_ctas, cudaStream_t stream) {
  const size_t smem = vfo_tile_smem_bytes(p.n_alt);
  auto kern = vfo_rowwarp_kernel<MODE, LITERAL>;
  cudaError_t e = grant_dynamic_smem((const void*)kern, 9 + MODE * 2 + (LITERAL ? 1 : 0), smem);
  if (e != cudaSuccess) return e;
  if (!p.use_pdl) {
    kern<<<(unsigned)n_ctas, kTileThreads, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_ctas);
  cfg.blockDim = dim3(kTileThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_vfo_rowwarp(const VfoParams& p, int mode, bool literal, int64_t n_ctas, cudaStream_t stream) {
  if (mode == 0)
    return literal ? launch_rowwarp_t<0, true>(p, n_ctas, stream) : launch_rowwarp_t<0, false>(p, n_ctas, stream);
  return literal ? launch_rowwarp_t<1, true>(p, n_ctas, stream) : launch_rowwarp_t<1, false>(p, n_ctas, stream);
}

int vfo_tile_ctas_per_sm(int n_alt, int max_smem_per_sm, bool solo_kernel) {
  const size_t a = sizeof(double) * 5 * (size_t)n_alt, b = vfo_tile_smem_bytes(n_alt);
  const size_t per_cta = (solo_kernel && a > b) ? a : b;
  const int by_smem = (int)((size_t)max_smem_per_sm / (per_cta + 1024));
  const int by_regs = solo_kernel ? kSoloMinBlocks : kTileMinBlocks;
  return by_smem < by_regs ? (by_smem < 1 ? 1 : by_smem) : by_regs;
}

template <int MODE, bool LITERAL>
static cudaError_t launch_tiles(const VfoParams& p, int64_t n_tiles, cudaStream_t stream) {
  const size_t smem = vfo_tile_smem_bytes(p.n_alt);
  auto kern = vfo_tile_kernel<MODE, LITERAL>;
  cudaError_t e = grant_dynamic_smem((const void*)kern, 1 + MODE * 2 + (LITERAL ? 1 : 0), smem);
  if (e != cudaSuccess) return e;
  if (!p.use_pdl) {
    kern<<<(unsigned)n_tiles, kTileThreads, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_tiles);
  cfg.blockDim = dim3(kTileThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_vfo_tiles(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream) {
  if (mode == 0) return literal ? launch_tiles<0, true>(p, n_tiles, stream) : launch_tiles<0, false>(p, n_tiles, stream);
  return literal ? launch_tiles<1, true>(p, n_tiles, stream) : launch_tiles<1, false>(p, n_tiles, stream);
}

int vfo_queue_ctas_per_sm() { return kQueueMinBlocks; }
int vfo_queue_threads() { return kQueueThreads; }
// levels of node buffer each of the eight CTAs of an SM can have (1 KB per CTA is reserved by the driver, ~0.5 KB static)
int vfo_queue_cap_nodes(int n_alt, int max_smem_per_sm) {
  const int by_smem = (int)(((size_t)max_smem_per_sm / kQueueMinBlocks - 1024 - 512) / sizeof(Node));
  return by_smem < n_alt ? (by_smem < 0 ? 0 : by_smem) : n_alt;
}

template <int MODE, bool LITERAL>
static cudaError_t launch_queue_t(const VfoParams& p, int64_t n_ctas, cudaStream_t stream) {
  const size_t smem = sizeof(Node) * (size_t)p.queue_cap_nodes;
  auto kern = vfo_queue_kernel<MODE, LITERAL>;
  cudaError_t e = grant_dynamic_smem((const void*)kern, 5 + MODE * 2 + (LITERAL ? 1 : 0), smem);
  if (e != cudaSuccess) return e;
  if (!p.use_pdl) {
    kern<<<(unsigned)n_ctas, kQueueThreads, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_ctas);
  cfg.blockDim = dim3(kQueueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_vfo_queue(const VfoParams& p, int mode, bool literal, int64_t n_ctas, cudaStream_t stream) {
  if (mode == 0) return literal ? launch_queue_t<0, true>(p, n_ctas, stream) : launch_queue_t<0, false>(p, n_ctas, stream);
  return literal ? launch_queue_t<1, true>(p, n_ctas, stream) : launch_queue_t<1, false>(p, n_ctas, stream);
}

template <int MODE, bool LITERAL>
static cudaError_t launch_solo_t(const VfoParams& p, int64_t n_tiles, cudaStream_t stream) {
  static_assert(kTileThreads == kThreads, "the solo kernel runs the row setup with the tile kernel's block size");
  const size_t a = sizeof(double) * 5 * (size_t)p.n_alt, b = vfo_tile_smem_bytes(p.n_alt);
  const size_t smem = a > b ? a : b;
  auto kern = vfo_solo_kernel<MODE, LITERAL>;
  cudaError_t e = grant_dynamic_smem((const void*)kern, 13 + MODE * 2 + (LITERAL ? 1 : 0), smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)n_tiles, kTileThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_vfo_solo(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream) {
  if (mode == 0)
    return literal ? launch_solo_t<0, true>(p, n_tiles, stream) : launch_solo_t<0, false>(p, n_tiles, stream);
  return literal ? launch_solo_t<1, true>(p, n_tiles, stream) : launch_solo_t<1, false>(p, n_tiles, stream);
}

cudaError_t launch_grid_multiplier(int n, size_t n_padded, double* m, double* dm, double* e, cudaStream_t stream) {
  const double step = (n > 1) ? 1.0 / (double)(n - 1) : 0.0;
  const int np = (int)n_padded;
  grid_multiplier_kernel<<<(np + 255) / 256, 256, 0, stream>>>(n, np, step, m, e);
  if (dm) grid_dmult_kernel<<<(np + 255) / 256, 256, 0, stream>>>(n, np, m, dm);
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

cudaError_t launch_residual(const double* vh, const double* vh_obs, int64_t n_profiles, int n_freq, double* residual,
                            double* chi2, cudaStream_t stream) {
  if (n_profiles <= 0) return cudaSuccess;
  int64_t blocks = (n_profiles + 7) / 8;
  if (blocks > 148 * 32) blocks = 148 * 32;
  residual_kernel<<<(unsigned)blocks, 256, 0, stream>>>(vh, vh_obs, n_profiles, n_freq, residual, chi2);
  return cudaGetLastError();
}

cudaError_t launch_argmin(const double* v, int64_t n, double* out2, cudaStream_t stream) {
  argmin_kernel<<<1, 1024, 0, stream>>>(v, n, out2);
  return cudaGetLastError();
}

cudaError_t launch_dfma_probe(double* out, int blocks, int iters, cudaStream_t stream) {
  dfma_probe_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0);
  return cudaGetLastError();
}

cudaError_t launch_math_selftest(int n, double* err2, cudaStream_t stream) {
  math_selftest_kernel<<<148 * 4, 256, 0, stream>>>(n, err2);
  return cudaGetLastError();
}

}  // namespace prhf

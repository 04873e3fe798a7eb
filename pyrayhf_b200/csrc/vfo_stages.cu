// Standalone stages of the vertical-forward-operator path (sm_100a, FP64).
//
// The product path (vfo_kernels.cu) fuses all of these into one pass and never materialises their
// [n_freq x n_points] arrays.  The reference exposes every stage as a public function, and its tutorial
// notebook plots the regridded arrays, so the same stages are offered here as separate operators with the
// reference's arithmetic order (PyRayHF/library.py, "lib"):
//   lib:75-97    den2freq                  -> den2freq_kernel
//   lib:120-137  find_X                    -> find_x_kernel
//   lib:140-158  find_Y                    -> find_y_kernel
//   lib:296-321  smooth_nonuniform_grid    -> smooth_grid_kernel
//   lib:324-438  regrid_to_nonuniform_grid -> row-setup kernel of vfo_kernels.cu (h_c) + regrid_write_kernel
//   lib:259-293  find_vh                   -> absmax_kernel (whole-array unmagnetised switch) + find_vh_kernel
// All of them are bound by HBM traffic (one read or write per element), not by arithmetic.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "vfo_device.cuh"
#include "vfo_kernels.h"

namespace prhf {

namespace {

constexpr int kStageThreads = 256;

inline unsigned stage_blocks(int64_t n, int per_thread = 1) {
  int64_t b = (n + (int64_t)kStageThreads * per_thread - 1) / ((int64_t)kStageThreads * per_thread);
  const int64_t cap = 148 * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ---- lib:93-96: any(density < 0) -> flag (the host raises ValueError), sqrt(density) * cp ----
__global__ void den2freq_kernel(const double* __restrict__ den, int64_t n, double* __restrict__ out,
                                int* __restrict__ negative_flag) {
  bool neg = false;
#pragma unroll 4
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = den[i];
    neg |= (v < 0.0);
    out[i] = __dmul_rn(__dsqrt_rn(v), kCp);
  }
  if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0 && negative_flag) atomicOr(negative_flag, 1);
}

// ---- lib:136: (sqrt(n) cp)^2 / f^2, strides 0 (scalar) or 1 ----
__global__ void find_x_kernel(const double* __restrict__ den, int64_t den_stride, const double* __restrict__ f_hz,
                              int64_t f_stride, int64_t n, double* __restrict__ X, int* __restrict__ negative_flag) {
  bool neg = false;
#pragma unroll 4
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = den[i * den_stride];
    neg |= (v < 0.0);
    X[i] = x_literal(v, f_hz[i * f_stride]);
  }
  if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0 && negative_flag) atomicOr(negative_flag, 1);
}

// ---- lib:157: g_p * b / f ----
__global__ void find_y_kernel(const double* __restrict__ f_hz, int64_t f_stride, const double* __restrict__ b,
                              int64_t b_stride, int64_t n, double* __restrict__ Y) {
#pragma unroll 4
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    Y[i] = y_literal(b[i * b_stride], f_hz[i * f_stride]);
}

// ---- lib:314-320 with general (start, end, sharpness) ----
__global__ void smooth_grid_kernel(int n, double step, double start, double end, double sharp, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u = __dmul_rn((double)i, step);                    // np.linspace(0, 1, n): arange(n) * step ...
  if (i == n - 1 && n > 1) u = 1.0;                         // ... endpoint forced
  if (n == 1) u = 0.0;
  const double fl = __dsub_rn(1.0, u);                      // lib:317
  const double factor = __ddiv_rn(__dsub_rn(exp(__dmul_rn(sharp, fl)), 1.0), __dsub_rn(exp(sharp), 1.0));   // lib:319
  x[i] = __dsub_rn(1.0, __dadd_rn(start, __dmul_rn(__dsub_rn(end, start), factor)));                         // lib:320
}

// ---- lib:410-427: stretched altitudes, their spacings, and the profile sampled on them ----
// grid = (chunks of points, frequency rows).  The truncated profile is staged in shared memory together with the
// per-level slopes (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]): numpy forms exactly that quotient for every query, so
// computing it once per level is bit-identical and takes three IEEE divisions per POINT out of the kernel, which
// would otherwise make it division-bound instead of write-bound.
struct RegridTable {
  const double* alt;     // [nt]
  const double* val[3];  // den, bmag, bpsi
  const double* slope[3];
};
__device__ __forceinline__ double regrid_interp(double x, int j, const RegridTable& t, int q, int n) {
  const double* fp = t.val[q];
  if (n == 1) return fp[0];                                 // numpy's single-node branch has no NaN test
  if (j == -2) return x;
  if (j == -1) return fp[0];
  if (j >= n - 1) return fp[n - 1];
  const double x0 = t.alt[j], f0 = fp[j];
  if (x0 == x) return f0;
  const double slope = t.slope[q][j];
  double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, x0)), f0);
  if (r != r) {                                             // numpy's NaN rescue
    const double f1 = fp[j + 1];
    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, t.alt[j + 1])), f1);
    if (r != r && f0 == f1) r = f0;
  }
  return r;
}
__global__ void __launch_bounds__(kStageThreads) regrid_write_kernel(const RegridParams p, const int n_alt) {
  extern __shared__ double s_tab[];
  // Launched with programmatic stream serialization behind the row-setup kernel: the table below depends on the
  // caller's profile only, so it is staged -- for ALL n_alt levels, the truncation index is not known yet -- while the
  // row setup is still running; what that kernel produces (the truncation index, the reflection heights) is read after
  // griddepcontrol.wait.  The slope of a level is a function of that level and the next one alone, so the first nt
  // entries are exactly what staging the truncated profile would give (slope[nt-1] is never used, regrid_interp).
  RegridTable t;
  double* w = s_tab;
  t.alt = w;
  for (int k = threadIdx.x; k < n_alt; k += blockDim.x) w[k] = p.alt[k];
  w += n_alt;
  const double* src[3] = {p.den, p.bmag, p.bpsi};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double* v = w;
    double* sl = w + n_alt;
    for (int k = threadIdx.x; k < n_alt; k += blockDim.x) {
      const double f0 = src[q][k];
      v[k] = f0;
      if (k + 1 < n_alt) sl[k] = __ddiv_rn(__dsub_rn(src[q][k + 1], f0), __dsub_rn(p.alt[k + 1], p.alt[k]));
    }
    t.val[q] = v;
    t.slope[q] = sl;
    w += 2 * n_alt;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");      // the row-setup grid has completed (no-op without PDL)
  const int nt = p.rec->nt;
  if (nt < 1) return;
  __syncthreads();
  const int r = blockIdx.y;
  const double alt0 = t.alt[0];
  const double span = __dsub_rn(p.row_hc[r], alt0);         // lib:413 (NaN on rows that never reflect)
  const int n = p.n_points;
  const int64_t base = (int64_t)r * n;
  const double inv_step = (nt > 1) ? (double)(nt - 1) / (t.alt[nt - 1] - alt0) : 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double h = __dadd_rn(__dmul_rn(p.mult[i], span), alt0);
    if (p.alt_out) p.alt_out[base + i] = h;
    if (p.dist_out) {
      const double hn = __dadd_rn(__dmul_rn(p.mult[i + 1], span), alt0);
      p.dist_out[base + i] = (i + 1 < n) ? __dsub_rn(hn, h) : kBackoff;                   // lib:415-416
    }
    // bracket: the uniform-grid guess when it verifies, else the binary search
    int j;
    {
      const int g = min(max(__double2int_rd((h - alt0) * inv_step), 0), nt - 1);
      if (h == h && t.alt[g] <= h && (g == nt - 1 ? h <= t.alt[g] : h < t.alt[g + 1])) j = g;
      else j = np_bracket(h, t.alt, nt);
    }
    if (p.den_out) p.den_out[base + i] = regrid_interp(h, j, t, 0, nt);                 // lib:424-426
    if (p.bmag_out) p.bmag_out[base + i] = regrid_interp(h, j, t, 1, nt);
    if (p.bpsi_out) p.bpsi_out[base + i] = regrid_interp(h, j, t, 2, nt);
  }
}

// ---- lib:201: nanmax(|Y|) over the whole array, as max over the IEEE bit patterns + 1 (0 = no finite-or-inf value) ----
__global__ void absmax_kernel(const double* __restrict__ y, int64_t n, unsigned long long* __restrict__ word) {
  unsigned long long m = 0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = fabs(y[i]);
    if (v == v) {
      const unsigned long long b = (unsigned long long)__double_as_longlong(v) + 1ull;
      m = b > m ? b : m;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(word, m);
}

// ---- lib:285-292: mu' from (X, Y, psi), nansum(mu' dh) per row, 0 -> NaN, + alt_min ----
// One CTA per row (n_cols > 1024) or one warp per row.
template <int MODE, bool LITERAL>
__global__ void __launch_bounds__(kStageThreads) find_vh_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                                const double* __restrict__ psi,
                                                                const double* __restrict__ dh, int64_t n_rows,
                                                                int64_t n_cols, double alt_min, int warp_rows,
                                                                const unsigned long long* __restrict__ word,
                                                                double* __restrict__ vh) {
  __shared__ double s_part[kStageThreads / 32];
  const unsigned long long w = *word;
  const bool iso = (w != 0ull) && (__longlong_as_double((long long)(w - 1ull)) < kYTol);   // lib:201
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t rows_per_cta = warp_rows ? (kStageThreads / 32) : 1;
  for (int64_t r0 = blockIdx.x * rows_per_cta; r0 < n_rows; r0 += (int64_t)gridDim.x * rows_per_cta) {
    const int64_t r = warp_rows ? r0 + wid : r0;
    double acc = 0.0;
    if (r < n_rows) {
      const int64_t base = r * n_cols;
      const int first = warp_rows ? lane : threadIdx.x;
      const int stride = warp_rows ? 32 : kStageThreads;
      for (int64_t c = first; c < n_cols; c += stride) {
        const double x = X[base + c];
        double mup;
        if (iso) {
          mup = iso_mup(x, nullptr);
        } else if (LITERAL) {
          mup = ah_literal<MODE>(x, Y[base + c], psi[base + c], nullptr);
        } else {
          double sn, cs;
          sincos(psi[base + c] * kDeg2Rad, &sn, &cs);
          mup = ah_fast<MODE>(x, Y[base + c], sn, cs, nullptr);
        }
        const double t = mup * dh[base + c];
        acc += (t == t) ? t : 0.0;                            // nansum (lib:288)
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (warp_rows) {
      if (lane == 0 && r < n_rows) vh[r] = (acc == 0.0 ? CUDART_NAN : acc) + alt_min;     // lib:290-292
    } else {
      if (lane == 0) s_part[wid] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < kStageThreads / 32; ++k) s += s_part[k];
        vh[r] = (s == 0.0 ? CUDART_NAN : s) + alt_min;
      }
      __syncthreads();
    }
  }
}

// ---- synthetic inputs on the device (SURVEY 8d "Common inputs"; pyrayhf_b200/synth.py is the host form) ----
// Two Chapman layers + centred axial dipole.  params[p] = {foF2 MHz, hmF2 km, H km, foE MHz, latitude deg}.
// One CTA per profile; replaces the PyIRI / IGRF input builders of lib:2390-2694, which are not available offline.
__device__ __forceinline__ double chapman_layer(double alt, double nm, double hm, double scale_h) {
  const double z = (alt - hm) / scale_h;
  return nm * exp(0.5 * (1.0 - z - exp(-z)));
}
__global__ void __launch_bounds__(kStageThreads) synth_profiles_kernel(const double* __restrict__ params, int64_t n_profiles,
                                                                       const double* __restrict__ alt, int n_alt,
                                                                       double* __restrict__ den, double* __restrict__ bmag,
                                                                       double* __restrict__ bpsi) {
  constexpr double kB0 = 3.12e-5, kRe = 6371.0;
  for (int64_t pr = blockIdx.x; pr < n_profiles; pr += gridDim.x) {
    const double* q = params + pr * 5;
    const double f2 = q[0] * 1e6 / kCp, fe = q[3] * 1e6 / kCp;
    const double nmf2 = f2 * f2, nme = fe * fe, hm = q[1], sh = q[2];
    const double lat = q[4] * kDeg2Rad;
    double sl, cl;
    sincos(lat, &sl, &cl);
    const double lat_factor = sqrt(1.0 + 3.0 * sl * sl);
    const double incl = atan2(2.0 * sl, cl) * (180.0 / CUDART_PI);
    const double psi = 90.0 - fabs(incl);
    for (int k = threadIdx.x; k < n_alt; k += blockDim.x) {
      const double a = alt[k];
      const int64_t o = pr * n_alt + k;
      den[o] = chapman_layer(a, nmf2, hm, sh) + chapman_layer(a, nme, 110.0, 8.0);
      const double r = kRe / (kRe + a);
      bmag[o] = kB0 * (r * r * r) * lat_factor;
      bpsi[o] = psi;
    }
  }
}

}  // namespace

cudaError_t launch_den2freq(const double* den, int64_t n, double* out, int* negative_flag, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  den2freq_kernel<<<stage_blocks(n), kStageThreads, 0, stream>>>(den, n, out, negative_flag);
  return cudaGetLastError();
}

cudaError_t launch_find_x(const double* den, int64_t den_stride, const double* f_hz, int64_t f_stride, int64_t n,
                          double* X, int* negative_flag, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  find_x_kernel<<<stage_blocks(n), kStageThreads, 0, stream>>>(den, den_stride, f_hz, f_stride, n, X, negative_flag);
  return cudaGetLastError();
}

cudaError_t launch_find_y(const double* f_hz, int64_t f_stride, const double* b, int64_t b_stride, int64_t n, double* Y,
                          cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  find_y_kernel<<<stage_blocks(n), kStageThreads, 0, stream>>>(f_hz, f_stride, b, b_stride, n, Y);
  return cudaGetLastError();
}

cudaError_t launch_smooth_grid(double start, double end, int n_points, double sharpness, double* x, cudaStream_t stream) {
  if (n_points <= 0) return cudaSuccess;
  const double step = (n_points > 1) ? 1.0 / (double)(n_points - 1) : 0.0;
  smooth_grid_kernel<<<(n_points + kStageThreads - 1) / kStageThreads, kStageThreads, 0, stream>>>(n_points, step, start,
                                                                                                  end, sharpness, x);
  return cudaGetLastError();
}

cudaError_t launch_regrid_write(const RegridParams& p, int n_alt, cudaStream_t stream) {
  if (p.n_freq <= 0 || p.n_points <= 0) return cudaSuccess;
  const size_t smem = sizeof(double) * 7 * (size_t)n_alt;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void*)regrid_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  // every CTA builds the slope table once (a few divisions per thread), so give it at least 16 points per thread,
  // unless that would leave fewer than ~4 CTAs per SM
  int per_thread = 16;
  while (per_thread > 2 && (int64_t)p.n_freq * ((p.n_points + kStageThreads * per_thread - 1) / (kStageThreads * per_thread)) < 148 * 4)
    per_thread >>= 1;
  int chunks = (p.n_points + kStageThreads * per_thread - 1) / (kStageThreads * per_thread);
  if (chunks < 1) chunks = 1;
  if (p.n_freq > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)chunks, (unsigned)p.n_freq);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kStageThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, regrid_write_kernel, p, n_alt);
}

cudaError_t launch_find_vh(const double* X, const double* Y, const double* psi, const double* dh, int64_t n_rows,
                           int64_t n_cols, double alt_min, int mode, bool literal, unsigned long long* scratch_word,
                           double* vh, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(scratch_word, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  if (n_cols > 0) absmax_kernel<<<stage_blocks(n_rows * n_cols, 4), kStageThreads, 0, stream>>>(Y, n_rows * n_cols, scratch_word);
  const int warp_rows = (n_cols <= 1024) ? 1 : 0;
  int64_t ctas = warp_rows ? (n_rows + kStageThreads / 32 - 1) / (kStageThreads / 32) : n_rows;
  if (ctas > 148 * 64) ctas = 148 * 64;
#define PRHF_LAUNCH(M, L) \
  find_vh_kernel<M, L><<<(unsigned)ctas, kStageThreads, 0, stream>>>(X, Y, psi, dh, n_rows, n_cols, alt_min, warp_rows, scratch_word, vh)
  if (mode == 0 && literal) PRHF_LAUNCH(0, true);
  else if (mode == 0) PRHF_LAUNCH(0, false);
  else if (literal) PRHF_LAUNCH(1, true);
  else PRHF_LAUNCH(1, false);
#undef PRHF_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_synth_profiles(const double* params, int64_t n_profiles, const double* alt, int n_alt, double* den,
                                  double* bmag, double* bpsi, cudaStream_t stream) {
  if (n_profiles <= 0 || n_alt <= 0) return cudaSuccess;
  int64_t ctas = n_profiles;
  if (ctas > 148 * 64) ctas = 148 * 64;
  synth_profiles_kernel<<<(unsigned)ctas, kStageThreads, 0, stream>>>(params, n_profiles, alt, n_alt, den, bmag, bpsi);
  return cudaGetLastError();
}

}  // namespace prhf

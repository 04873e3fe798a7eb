// Device-side math of the vertical forward operator (sm_100a, FP64).
//
// Reference behaviour being replaced (PyRayHF/library.py, "lib"):
//   lib:120-158  find_X / find_Y            -> x_literal / y_literal and the per-row scale factors
//   lib:161-256  find_mu_mup                -> ah_literal (operation by operation), ah_fast (sign-safe), ah_hot (hot loop)
//   lib:424-426  np.interp of den/bmag/bpsi -> point_term in vfo_kernels.cu (numpy arr_interp semantics) / fma form
//
// Tensor cores are not used: the path is ~50 dependent FP64 scalar operations per grid point with
// no contraction structure; it is bounded by the FP64 pipe (DESIGN.md "Roofline").
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace prhf {

constexpr double kCp = 8.97866275;          // lib:61
constexpr double kGp = 2.799249247e10;      // lib:64
constexpr double kSharp = 10.0;             // lib:363
// lib:314-320 with sharpness 10: m = 1 - (E - 1) / (e^10 - 1) = kStretchA - kStretchB * E, E = exp(10 (1 - u))
constexpr double kStretchDen = 22025.465794806718;    // fl(fl(e^10) - 1), the reference's denominator
constexpr double kStretchB = 4.5401991009687765e-05;  // fl(1 / kStretchDen)
constexpr double kStretchA = 1.0000454019910097;      // fl(1 + kStretchB)
constexpr double kBackoff = 1e-6;           // lib:378 (the dh argument is overwritten)
constexpr double kYTol = 1e-12;             // lib:163
constexpr double kDeg2Rad = 0.017453292519943295;  // fl(pi/180): np.deg2rad(x) == x * (pi/180)
constexpr double kScreenTol = 6e-15;        // K1 node screening: > 4x the 13-ulp bound on |screen - literal| / (|X|+|Y|)

// ---- lib:136 and lib:157 with the reference's rounding order (no contraction) ----
__device__ __forceinline__ double x_literal(double den, double f_hz) {
  const double fp = __dmul_rn(__dsqrt_rn(den), kCp);
  return __ddiv_rn(__dmul_rn(fp, fp), __dmul_rn(f_hz, f_hz));
}
__device__ __forceinline__ double y_literal(double b, double f_hz) {
  return __ddiv_rn(__dmul_rn(kGp, b), f_hz);
}

// ---- lib:209-254, every operation in the reference's order (IEEE div / sqrt, libdevice sincos) ----
// Returns mu' ; *mu_out (optional) receives mu.  NaN where the reference produces NaN.
template <int MODE>
__device__ __forceinline__ double ah_literal(double X, double Y, double psi_deg, double* mu_out) {
  const double sgn = (MODE == 0) ? 1.0 : -1.0;
  const double rad = __dmul_rn(psi_deg, kDeg2Rad);
  double s, c;
  sincos(rad, &s, &c);
  const double YT = __dmul_rn(Y, s), YL = __dmul_rn(Y, c);
  const double Xm1 = __dsub_rn(1.0, X);
  const double YT2 = __dmul_rn(YT, YT);
  const double YL2 = __dmul_rn(YL, YL);
  const double Xm12 = __dmul_rn(Xm1, Xm1);
  const double alpha = __dadd_rn(__dmul_rn(0.25, __dmul_rn(YT2, YT2)), __dmul_rn(YL2, Xm12));
  const double beta = __dsqrt_rn(alpha);
  const double D = __dadd_rn(__dsub_rn(Xm1, __dmul_rn(0.5, YT2)), __dmul_rn(sgn, beta));
  const double XXm1 = __dmul_rn(X, Xm1);
  double u = __dsub_rn(1.0, __ddiv_rn(XXm1, D));
  if (u < 0.0) u = CUDART_NAN;
  double mu = __dsqrt_rn(u);
  if (mu > 1.0) mu = CUDART_NAN;
  const double dbdx = __ddiv_rn(__dmul_rn(-YL2, Xm1), beta);
  const double dDdX = __dadd_rn(-1.0, __dmul_rn(sgn, dbdx));
  const double dady = __dadd_rn(__dmul_rn(__dmul_rn(YT2, YT), s),
                                __dmul_rn(__dmul_rn(__dmul_rn(2.0, YL), Xm12), c));
  const double dbdy = __ddiv_rn(__dmul_rn(0.5, dady), beta);
  const double dDdY = __dadd_rn(__dmul_rn(-YT, s), __dmul_rn(sgn, dbdy));
  const double dmudY = __ddiv_rn(__dmul_rn(XXm1, dDdY), __dmul_rn(__dmul_rn(2.0, mu), __dmul_rn(D, D)));
  const double dmudX = __dmul_rn(__ddiv_rn(1.0, __dmul_rn(__dmul_rn(2.0, mu), D)),
                                 __dadd_rn(__dsub_rn(__dmul_rn(2.0, X), 1.0),
                                           __dmul_rn(__ddiv_rn(XXm1, D), dDdX)));
  if (mu_out) *mu_out = mu;
  return __dsub_rn(mu, __dadd_rn(__dmul_rn(__dmul_rn(2.0, X), dmudX), __dmul_rn(Y, dmudY)));
}

// ---- lib:202-206 (isotropic branch) ----
__device__ __forceinline__ double iso_mup(double X, double* mu_out) {
  const double mu2 = 1.0 - X;
  const double mu = (mu2 > 0.0) ? sqrt(mu2) : CUDART_NAN;
  if (mu_out) *mu_out = mu;
  return (isfinite(mu) && mu > 0.0) ? 1.0 / mu : CUDART_NAN;
}

// ---- reciprocal and reciprocal square root without the libdevice special-case paths ----
// MUFU.RCP64H / MUFU.RSQ64H seed (rcp/rsqrt.approx.ftz.f64, measured max relative error 9.9e-7 / 9.2e-7
// on B200) followed by ONE cubically convergent step: 3 (rcp) / 5 (rsqrt) FP64 instructions, no branches.
// Measured max relative error against IEEE division / sqrt (prhf_selftest_math, 2^24 samples over
// 1e-30..1e30): 2.2e-16 / 2.7e-16.  Zero, negative, infinite and denormal inputs yield NaN/inf rather than
// the IEEE special values; on this path they only occur at points whose term the reference drops as NaN
// anyway (lib:233, lib:288).
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);                      // y (1 + e + e^2)
}
// (3/8 comes from the constant bank: fma(e, 0.375, 0.5) holds two literals and only one fits the instruction, so the
//  compiler would otherwise rebuild 0.375 in a register pair inside every loop that is short of registers.)
static __constant__ double kThreeEighths = 0.375;
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);              // 1 - x y^2
  // y (1 + e/2 + 3 e^2 / 8) as y + (y e)(1/2 + 3 e / 8): the two factors are independent of each other, so the result
  // is four dependent FP64 operations behind the seed instead of five (the grid loop is a latency chain)
  return fma(y * e, fma(e, kThreeEighths, 0.5), y);
}
// 1/sqrt(x) and sqrt(x) from the same seed and the same correction, both five dependent operations behind the seed
// (sqrt(x) as x * rsqrt_fast(x) would be a sixth).  s = x y is ~sqrt(x); both are scaled by 1 + e/2 + 3 e^2/8.
__device__ __forceinline__ double rsqrt_sqrt_fast(double x, double* sqrt_out) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double s = x * y;
  const double e = fma(-s, y, 1.0);
  const double c = e * fma(e, kThreeEighths, 0.5);
  *sqrt_out = fma(s, c, s);
  return fma(y, c, y);
}

// ---- per-row scale factors of the screen and the fast paths: X = den * kx, Y = b * ky ----
// (never used where the reference's rounding order matters: those places call x_literal / y_literal.)
// Reciprocals through rcp_fast instead of IEEE division: a division is ~400 cycles of pure latency and the row
// prologue used to have six of them back to back on its critical path.  Relative error <= 5e-16.
__device__ __forceinline__ void row_scales(double f_hz, double* kx, double* ky) {
  const double af = fabs(f_hz);
  *kx = (kCp * kCp) * rcp_fast(af * af);
  *ky = copysign(kGp * rcp_fast(af), f_hz);
}

// ---- restructured Appleton-Hartree: same formulas as lib:209-254, algebraically rearranged ----
//   a = YT^2/2, w = YL^2 Xm1, alpha = a^2 + w Xm1, beta = sqrt(alpha), P = a + beta
//   X-mode:  D = Xm1 - P                      (no cancellation: D -> Y(1-Y) at reflection)
//   O-mode:  D = Xm1 (1 + g), g = w / P       (cancellation-free form of Xm1 - a + beta)
//   Y dD/dY = -2a + s (beta + a^2/beta),  dD/dX = -1 - s w / beta      (s = +1 O, -1 X)
//   mu' = mu - [ X (2X - 1 + q dD/dX) + q (Y dD/dY) / 2 ] / (mu D),   q = X Xm1 / D
// One reciprocal and two reciprocal square roots per point instead of 7 divides + 2 square roots.
// Agreement with a long-double evaluation of lib:209-254: <= 3e-12 (tests/test_oracle_*).
template <int MODE>
__device__ __forceinline__ double ah_fast(double X, double Y, double sn, double cs, double* mu_out) {
  const double YT = Y * sn, YL = Y * cs;
  const double Xm1 = 1.0 - X;
  const double a = 0.5 * (YT * YT);
  const double w = (YL * YL) * Xm1;
  const double a2 = a * a;
  const double alpha = fma(w, Xm1, a2);
  const double rb = rsqrt_fast(alpha);
  const double beta = alpha * rb;
  const double P = a + beta;
  double invD, q, u, dDdX, YdDdY;
  if (MODE == 1) {
    const double D = Xm1 - P;
    invD = rcp_fast(D);
    q = (X * Xm1) * invD;
    u = 1.0 - q;
    dDdX = fma(w, rb, -1.0);
    YdDdY = -fma(a2, rb, beta) - 2.0 * a;
  } else {
    const double R = rcp_fast(Xm1 * (P + w));
    const double XR = Xm1 * R;
    invD = P * R;
    q = (X * P) * XR;
    u = fma(Xm1, P, w) * XR;
    dDdX = -fma(w, rb, 1.0);
    YdDdY = fma(a2, rb, beta) - 2.0 * a;
  }
  const double rmu = rsqrt_fast(u);
  const double mu = fmin(u * rmu, 1.0);
  const double br = fma(0.5 * q, YdDdY, X * fma(q, dDdX, fma(2.0, X, -1.0)));
  double mup = fma(-(invD * rmu), br, mu);
  // lib:233: u < 0 -> NaN.  lib:238: mu > 1 -> NaN, where the reference's mu is sqrt(fl(1 - q)): that exceeds 1
  // exactly when fl(1 - q) >= 1 + 2^-51 (the square root of 1 + 2^-52 rounds back to 1).  The same rounding is
  // applied to q here, so that vacuum (q == 0, kept) and the faintest plasma below the gyrofrequency in X-mode
  // (q = -1e-15, dropped) come out as in the reference; u itself carries a reciprocal's rounding in O-mode.
  const bool ok = (u >= 0.0) && ((1.0 - q) <= 1.0000000000000002);
  if (!ok) mup = CUDART_NAN;
  if (mu_out) {
    double mu_ret = ok ? mu : CUDART_NAN;
    if (alpha == 0.0) {
      // beta == 0 (a level without field, or Y_T == 0 at X == 1): the reference's derivative terms are 0/0, so
      // mu' is NaN, but its mu = sqrt(1 - X (1-X) / (1-X)) is finite and the Snell tracers keep such a level
      // (lib:229-238 with beta = 0).  The reciprocal square root above turned everything into NaN.
      const double uz = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(X, Xm1), Xm1));
      mu_ret = (uz >= 0.0 && uz <= 1.0) ? __dsqrt_rn(uz) : CUDART_NAN;
    }
    *mu_out = mu_ret;
  }
  return mup;
}

// ---- hot-loop form: the reciprocal of D folded into the second reciprocal square root ----
// Used for every grid point of a row that reflects above the first level.
//   X-mode:  E = D - X Xm1 (= D mu^2), rs = 1/sqrt(D E):  mu = |E rs|, 1/D = (E rs) rs,
//            1/(D mu) = copysign(rs, E rs)           -- exact for either sign of D
//   O-mode:  N = Xm1 P + w, G = P + w, mu^2 = N/G, rs = 1/sqrt(Xm1^2 N G):  mu = Xm1 N rs,
//            1/G = (Xm1 N rs)(Xm1 rs), q = X P / G, 1/(D mu) = P rs   -- assumes Xm1 > 0, which holds at
//            every grid point below the X = 1 reflection level (rows that start above it take ah_fast)
// Inputs: YTh = Y sin(psi) / sqrt(2) (so that a = YTh^2), YL = Y cos(psi).  Returns mu', mu and q = X (1-X) / D; the
// caller tests validity (keep_term) on the bit patterns with integer instructions, keeping compares off the
// FP64 pipe.  36 (X) / 37 (O) FP64 instructions.
template <int MODE>
__device__ __forceinline__ double ah_hot(double X, double YTh, double YL, double* mu_out, double* q_out) {
  const double Xm1 = 1.0 - X;
  const double a = YTh * YTh;
  const double w = (YL * YL) * Xm1;
  const double a2 = a * a;
  const double alpha = fma(w, Xm1, a2);
  double beta;
  const double rb = rsqrt_sqrt_fast(alpha, &beta);
  const double P = a + beta;
  const double T = fma(a2, rb, beta);                   // beta + a^2 / beta
  double mu, nc, q, dDdX, hYd;                          // nc = -1/(D mu), hYd = (Y dD/dY) / 2
  if (MODE == 1) {
    const double D = Xm1 - P;
    const double XX = X * Xm1;
    const double E = D - XX;
    const double rs = rsqrt_fast(D * E);
    const double t1 = E * rs;
    mu = fabs(t1);
    q = XX * (t1 * rs);
    // nc = -copysign(rs, t1) = -1/(D mu), sign bit assembled with ONE integer instruction (rs > 0 or NaN): written
    // as copysign + negation the compiler negates t1 on the FP64 pipe first
    nc = __hiloint2double(__double2hiint(rs) | (~__double2hiint(t1) & (int)0x80000000), __double2loint(rs));
    dDdX = fma(w, rb, -1.0);
    hYd = fma(-0.5, T, -a);
  } else {
    const double N = fma(Xm1, P, w);
    const double G = P + w;
    const double z = Xm1 * N;
    const double rs = rsqrt_fast(z * (Xm1 * G));
    const double v = z * rs;
    mu = fabs(v);
    q = (X * P) * (v * (Xm1 * rs));
    nc = -(P * rs);
    dDdX = -fma(w, rb, 1.0);
    hYd = fma(0.5, T, -a);
  }
  // 2X - 1 as X - (1 - X): one addition without a literal (fma(2.0, X, -1.0) needs 2.0 in a register pair)
  const double br = fma(q, hYd, X * fma(q, dDdX, X - Xm1));
  *mu_out = mu;
  *q_out = q;
  return fma(nc, br, mu);
}

// mu' is kept when the reference keeps it: mu not NaN (lib:233), mu <= 1 (lib:238), mu' itself not NaN
// (nansum, lib:288).  Integer tests on the IEEE bit patterns.
// lib:238 is decided on q = X (1-X) / D rather than on mu: the reference's mu is sqrt(fl(1 - q)), which exceeds 1
// exactly when q <= -3 * 2^-53 (fl(1 - q) >= 1 + 2^-51; the square root of 1 + 2^-52 rounds back to 1).  The mu
// computed here comes out of a reciprocal square root and is only good to 3e-16, which is not enough to tell
// vacuum (q == 0: kept, e.g. zero density below the layer) from the faintest plasma below the gyrofrequency in
// X-mode (q = -1e-15: dropped).  Both cases were found by tests/test_gpu_fuzz.py.
// As unsigned integers the doubles order as: +0 ... +inf, NaN, -0 ... -inf, and both thresholds have a zero low
// word, so each test is ONE 32-bit compare on the high word:
//   q > -3 * 2^-53          <=>  hi(q)  < 0xBCA80000
//   mu' finite               <=>  hi(|mu'|) < 0x7FF00000
// A NaN mu (lib:233: D E < 0 under the reciprocal square root) makes mu' NaN, so mu needs no test of its own; an
// infinite mu' cannot come out of ah_hot (mu == 0 gives 0 * inf = NaN there).
// acc += mu' * w when keep_term holds: the two integer tests feed the predicate of ONE predicated DFMA (as a select
// the compiler spends four FSEL per grid point on the two halves of the operand).
__device__ __forceinline__ void add_kept(double& acc, double mup, double q, double w) {
  asm("{\n\t"
      ".reg .pred k1, k2;\n\t"
      ".reg .u32 ph;\n\t"
      "and.b32 ph, %3, 0x7fffffff;\n\t"
      "setp.lt.u32 k1, ph, 0x7ff00000;\n\t"
      "setp.lt.and.u32 k2, %4, 0xBCA80000, k1;\n\t"
      "@k2 fma.rn.f64 %0, %1, %2, %0;\n\t"
      "}"
      : "+d"(acc)
      : "d"(mup), "d"(w), "r"(__double2hiint(mup)), "r"(__double2hiint(q)));
}
__device__ __forceinline__ bool keep_term(double mup, double q) {
  const unsigned ph = (unsigned)__double2hiint(mup) & 0x7fffffffu;
  const unsigned qh = (unsigned)__double2hiint(q);
  return (ph < 0x7ff00000u) && (qh < 0xBCA80000u);
}

// sin/cos of (r_k + delta) from the node's sin/cos and a short Taylor series in delta (|delta| <= 0.05:
// truncation < 6e-18).  Replaces a full-range sincos per grid point.
__device__ __forceinline__ void rotate_sincos(double sk, double ck, double delta, double* sn, double* cs) {
  const double d2 = delta * delta;
  double ps = fma(d2, -1.0 / 5040.0, 1.0 / 120.0);
  ps = fma(d2, ps, -1.0 / 6.0);
  ps = fma(d2, ps, 1.0);
  const double sd = delta * ps;                       // sin(delta)
  double pc = fma(d2, 1.0 / 40320.0, -1.0 / 720.0);
  pc = fma(d2, pc, 1.0 / 24.0);
  pc = fma(d2, pc, -0.5);
  const double cdm1 = d2 * pc;                        // cos(delta) - 1
  *sn = fma(ck, sd, fma(sk, cdm1, sk));
  *cs = fma(-sk, sd, fma(ck, cdm1, ck));
}
constexpr double kMaxRotateStep = 0.05;               // rad; larger per-segment steps use sincos()
// Second-order version for |delta| <= 4e-4 rad (0.023 deg per profile level; truncation delta^3/6 < 1.1e-11).
__device__ __forceinline__ void rotate_sincos_small(double sk, double ck, double delta, double* sn, double* cs) {
  const double e2 = (-0.5 * delta) * delta;
  *sn = fma(sk, e2, fma(delta, ck, sk));
  *cs = fma(ck, e2, fma(-delta, sk, ck));
}
constexpr double kSmallRotateStep = 4e-4;

// ---- numpy arr_interp for one query against the staged altitude axis (left = fp[0], right = fp[n-1]) ----
// Returns the bracket: -2 NaN query, -1 below the axis, n above it, else the last j with xp[j] <= x.
// The two range tests come first and in numpy's order (binary_search_with_guess: "key > arr[len - 1]" before
// "key < arr[0]"), which is what defines np.interp on a DEcreasing axis: every query above the last (smallest)
// coordinate takes the right-hand fill value fp[n-1], whatever lies in between.
__device__ __forceinline__ int np_bracket(double x, const double* __restrict__ xp, int n) {
  if (x != x) return -2;
  if (x > xp[n - 1]) return n;
  if (x < xp[0]) return -1;
  int lo = 0, hi = n - 1;                                   // xp[lo] <= x <= xp[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  return (xp[hi] <= x) ? hi : lo;
}
__device__ __forceinline__ double np_interp_at(double x, int j, const double* __restrict__ xp,
                                               const double* __restrict__ fp, int n) {
  if (n == 1) return fp[0];                                 // numpy's single-node branch has no NaN test
  if (j == -2) return x;
  if (j == -1) return fp[0];
  if (j >= n - 1) return fp[n - 1];
  const double x0 = xp[j], f0 = fp[j];
  if (x0 == x) return f0;
  const double x1 = xp[j + 1], f1 = fp[j + 1];
  const double slope = __ddiv_rn(__dsub_rn(f1, f0), __dsub_rn(x1, x0));
  double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, x0)), f0);
  if (r != r) {                                             // numpy's NaN rescue
    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, x1)), f1);
    if (r != r && f0 == f1) r = f0;
  }
  return r;
}

}  // namespace prhf

/*
 * TEST INFRASTRUCTURE ONLY -- scalar C restatement of PyRayHF's vertical forward operator.
 *
 * One (profile, frequency) row at a time, no [n_freq x n_points] arrays.  Follows
 * /root/reference/PyRayHF/library.py ("lib") line by line:
 *   lib:459-509 vertical_forward_operator, lib:324-438 regrid_to_nonuniform_grid,
 *   lib:296-321 smooth_nonuniform_grid, lib:120-158 find_X / find_Y,
 *   lib:161-256 find_mu_mup, lib:259-293 find_vh,
 * and the numpy primitives the reference executes through (numpy is an unpinned
 * dependency, pyproject.toml:42; semantics restated from numpy 2.3.5's published
 * compiled_base.c `arr_interp` / `binary_search_with_guess`, loops_utils.h
 * `pairwise_sum`, and `np.maximum` NaN propagation).
 *
 * Variants (argument `variant`):
 *   0  literal float64 -- every rounding in the reference's order (libm exp/pow/sin/cos
 *      stand in for numpy's SIMD versions, which differ by <= 2 ulp: not bit-identical,
 *      agreement ~1e-12 in X-mode).
 *   1  "truth" -- the same float64 h_i, dh_i, interpolants, X and Y as variant 0, then
 *      lib:209-254 and the sum in 80-bit long double with the cancellation-free
 *      O-mode denominator (hard part 0).  This is the value
 *      oracle for O-mode, where the float64 reference itself is only good to ~1e-5.
 *
 * Parity status: PINNED by tests/test_oracle_golden.py (fixtures generated from the live
 * reference by tests/make_golden.py) and tests/test_oracle_vs_reference.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline legs may load this.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared -fPIC; OpenMP is not in this image, so
 * host threads are plain pthreads pulling work items off an atomic counter).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define CP 8.97866275             /* lib:61 */
#define GP 2.799249247e10         /* lib:64 */
#define SHARP 10.0                /* lib:363 */
#define BACKOFF 1e-6              /* lib:378 */
#define YTOL 1e-12                /* lib:163 */

/* np.linspace(0,1,n) then lib:317-320.  m[n]. */
void vfo_oracle_multiplier(int n, double *m) {
  if (n <= 0) return;
  double step = (n > 1) ? 1.0 / (double)(n - 1) : 0.0;
  double den = exp(SHARP) - 1.0;
  for (int i = 0; i < n; ++i) {
    double u = (double)i * step;
    if (i == n - 1 && n > 1) u = 1.0;      /* linspace forces the endpoint */
    double fl = 1.0 - u;
    double factor = (exp(SHARP * fl) - 1.0) / den;
    m[i] = 1.0 - (0.0 + (1.0 - 0.0) * factor);
  }
}

/* numpy binary_search_with_guess result for a sorted table: -1, n, or last j with xp[j] <= x */
static int np_bracket(double x, const double *xp, int n) {
  if (x > xp[n - 1]) return n;
  if (x < xp[0]) return -1;
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if (x >= xp[mid]) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

/* numpy arr_interp for one query (left=fp[0], right=fp[n-1]) */
static double np_interp1(double x, const double *xp, const double *fp, int n) {
  if (n == 1) return fp[0];   /* numpy's lenxp == 1 branch has no NaN test: interp(NaN) -> fp[0] */
  if (isnan(x)) return x;
  int j = np_bracket(x, xp, n);
  if (j == -1) return fp[0];
  if (j == n) return fp[n - 1];
  if (j == n - 1) return fp[j];
  if (xp[j] == x) return fp[j];
  double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  double r = slope * (x - xp[j]) + fp[j];
  if (isnan(r)) {
    r = slope * (x - xp[j + 1]) + fp[j + 1];
    if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
  }
  return r;
}

/* numpy pairwise_sum (contiguous) */
static double np_pairwise(const double *a, long n) {
  if (n < 8) {
    double r = 0.0;
    for (long i = 0; i < n; ++i) r += a[i];
    return r;
  } else if (n <= 128) {
    double r[8];
    long i;
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
  }
}

static double x_of(double den, double f_hz) {      /* lib:96, lib:136 */
  double fp = sqrt(den) * CP;
  return (fp * fp) / (f_hz * f_hz);
}
static double y_of(double b, double f_hz) {        /* lib:157 */
  return (GP * b) / f_hz;
}

/* lib:209-254, float64, every operation in the reference's order */
static double mup_literal(double X, double Y, double psi, double sgn) {
  double rad = psi * (M_PI / 180.0);
  double s = sin(rad), c = cos(rad);
  double YT = Y * s, YL = Y * c;
  double Xm1 = 1.0 - X;
  double alpha = 0.25 * pow(YT, 4.0) + (YL * YL) * (Xm1 * Xm1);
  double beta = sqrt(alpha);
  double D = (Xm1 - 0.5 * (YT * YT)) + sgn * beta;
  double u = 1.0 - (X * Xm1) / D;
  if (u < 0) u = NAN;
  double mu = sqrt(u);
  if (mu < 0.0) mu = 0.0;
  if (mu > 1.0) mu = NAN;
  double dbdx = (-(YL * YL) * Xm1) / beta;
  double dDdX = -1.0 + sgn * dbdx;
  double dady = (pow(YT, 3.0) * s) + (((2.0 * YL) * (Xm1 * Xm1)) * c);
  double dbdy = (0.5 * dady) / beta;
  double dDdY = (-YT) * s + sgn * dbdy;
  double dmudY = ((X * Xm1) * dDdY) / ((2.0 * mu) * (D * D));
  double dmudX = (1.0 / ((2.0 * mu) * D)) * (((2.0 * X) - 1.0) + (((X * Xm1) / D) * dDdX));
  return mu - (((2.0 * X) * dmudX) + (Y * dmudY));
}

/* lib:202-206 */
static double mup_iso(double X) {
  double mu2 = 1.0 - X;
  double mu = (mu2 > 0.0) ? sqrt(mu2) : NAN;
  return (isfinite(mu) && mu > 0.0) ? 1.0 / mu : NAN;
}

/* lib:209-254 in long double from float64 X, Y, psi; O-mode D in the cancellation-free form
 * D = Xm1*(1+g), g = YL^2*Xm1/(beta + YT^2/2)  (algebraically identical to lib:229). */
static long double mup_truth(double Xd, double Yd, double psid, int mode) {
  long double X = Xd, Y = Yd;
  long double rad = (long double)psid * (M_PIl / 180.0L);
  long double s = sinl(rad), c = cosl(rad);
  long double YT = Y * s, YL = Y * c, Xm1 = 1.0L - X;
  long double hYT2 = 0.5L * YT * YT;
  long double beta = sqrtl(0.25L * YT * YT * YT * YT + YL * YL * Xm1 * Xm1);
  long double sgn, D, q, u;
  if (mode == 0) {
    sgn = 1.0L;
    long double g = YL * YL * Xm1 / (beta + hYT2);
    D = Xm1 * (1.0L + g);
    q = X / (1.0L + g);
    u = (Xm1 + g) / (1.0L + g);
  } else {
    sgn = -1.0L;
    D = Xm1 - hYT2 - beta;
    q = X * Xm1 / D;
    u = 1.0L - q;
  }
  if (u < 0) return NAN;
  long double mu = sqrtl(u);
  if (mu > 1.0L) return NAN;
  long double dbdx = -YL * YL * Xm1 / beta;
  long double dDdX = -1.0L + sgn * dbdx;
  long double dady = YT * YT * YT * s + 2.0L * YL * Xm1 * Xm1 * c;
  long double dbdy = 0.5L * dady / beta;
  long double dDdY = -YT * s + sgn * dbdy;
  long double dmudY = (X * Xm1 * dDdY) / (2.0L * mu * D * D);
  long double dmudX = (1.0L / (2.0L * mu * D)) * (2.0L * X - 1.0L + q * dDdX);
  return mu - (2.0L * X * dmudX + Y * dmudY);
}

typedef struct {
  double s_mag;   /* nansum with the magnetised formulas */
  double s_iso;   /* nansum with the isotropic formulas  */
  double ymax;    /* nanmax |Y| over this row's grid points (NaN if none) */
  double hc;      /* reflection height after back-off, NaN when the row is dead */
} row_out_t;

/* One (profile, frequency) row.  Tables are already truncated to [0, nt).  scratch: 2*n doubles. */
static void eval_row(double f_mhz, const double *alt, const double *den, const double *bmag,
                     const double *bpsi, int nt, int mode, int n, const double *m, int variant,
                     double *scratch, row_out_t *out) {
  double f_hz = f_mhz * 1e6;                               /* lib:491 */
  /* lib:380-399: critical curve with running max, validity from its last element */
  double *crit = (double *)malloc(sizeof(double) * (size_t)nt);
  double run = 0.0;
  for (int k = 0; k < nt; ++k) {
    double v = x_of(den[k], f_hz);
    if (mode == 1) v = v + y_of(bmag[k], f_hz);
    if (k == 0) run = v;
    else run = (run >= v || isnan(run)) ? run : v;         /* np.maximum */
    crit[k] = run;
  }
  int valid = crit[nt - 1] >= 1.0;
  double hc = np_interp1(1.0, crit, alt, nt);              /* lib:403-404 */
  hc = valid ? hc - BACKOFF : NAN;                         /* lib:407 */
  free(crit);
  out->hc = hc;

  double sgn = (mode == 0) ? 1.0 : -1.0;
  double *t_mag = scratch, *t_iso = scratch + n;
  long double acc_mag = 0.0L, acc_iso = 0.0L;
  double ymax = NAN;
  double span = hc - alt[0];
  double h = m[0] * span + alt[0];                         /* lib:413 */
  for (int i = 0; i < n; ++i) {
    double hn = (i + 1 < n) ? m[i + 1] * span + alt[0] : 0.0;
    double dh = (i + 1 < n) ? hn - h : BACKOFF;            /* lib:415-416 */
    double d_i = np_interp1(h, alt, den, nt);              /* lib:424-426 */
    double b_i = np_interp1(h, alt, bmag, nt);
    double p_i = np_interp1(h, alt, bpsi, nt);
    double X = x_of(d_i, f_hz);                            /* lib:500 */
    double Y = y_of(b_i, f_hz);                            /* lib:503 */
    double ay = fabs(Y);
    if (!isnan(ay) && (isnan(ymax) || ay > ymax)) ymax = ay;
    if (variant == 0) {
      double a = mup_literal(X, Y, p_i, sgn) * dh;         /* lib:288 */
      double b = mup_iso(X) * dh;
      t_mag[i] = isnan(a) ? 0.0 : a;                       /* nansum: NaN -> 0 */
      t_iso[i] = isnan(b) ? 0.0 : b;
    } else {
      long double a = mup_truth(X, Y, p_i, mode) * (long double)dh;
      long double X_l = X;
      long double mu2 = 1.0L - X_l;
      long double b = (mu2 > 0.0L) ? (long double)dh / sqrtl(mu2) : NAN;
      if (!isnan(a)) acc_mag += a;
      if (!isnan(b)) acc_iso += b;
    }
    h = hn;
  }
  if (variant == 0) {
    out->s_mag = np_pairwise(t_mag, n);
    out->s_iso = np_pairwise(t_iso, n);
  } else {
    out->s_mag = (double)acc_mag;
    out->s_iso = (double)acc_iso;
  }
  out->ymax = ymax;
}


/* ---- host threading: pthreads + an atomic work counter (no OpenMP runtime in this image) ---- */
static void run_threads(void *(*fn)(void *), void *arg, int n_threads, int64_t n_items) {
  if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (n_threads > n_items) n_threads = (int)(n_items > 0 ? n_items : 1);
  if (n_threads <= 1) { fn(arg); return; }
  pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int i = 0; i < n_threads; ++i) pthread_create(&t[i], NULL, fn, arg);
  for (int i = 0; i < n_threads; ++i) pthread_join(t[i], NULL);
  free(t);
}

typedef struct {
  const double *freq_mhz, *alt, *den, *bmag, *bpsi;
  int nt, mode, n;
  const double *m;
  int variant;
  row_out_t *rows;
  int n_freq;
  int64_t next;
} row_job_t;

static void *row_worker(void *p) {
  row_job_t *j = (row_job_t *)p;
  double *scratch = (double *)malloc(sizeof(double) * 2 * (size_t)(j->n > 0 ? j->n : 1));
  for (;;) {
    int64_t r = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (r >= j->n_freq) break;
    eval_row(j->freq_mhz[r], j->alt, j->den, j->bmag, j->bpsi, j->nt, j->mode, j->n, j->m, j->variant,
             scratch, &j->rows[r]);
  }
  free(scratch);
  return NULL;
}

/*
 * One profile, n_freq frequencies.  Returns 0 ok, 1 negative density below the peak
 * (lib:93-94), 2 peak at index 0 (IndexError at lib:399), 3 bad mode (lib:396).
 * multiplier may be NULL (computed with libm).  hc_out may be NULL.
 */
int vfo_oracle_profile(const double *freq_mhz, int n_freq, const double *den, const double *bmag,
                       const double *bpsi, const double *alt, int n_alt, int mode, int n_points,
                       const double *multiplier, int variant, int n_threads, double *vh_out,
                       double *hc_out) {
  if (mode != 0 && mode != 1) return 3;
  /* lib:371: argmax, first maximum, NaN wins */
  int kmax = 0;
  for (int k = 0; k < n_alt; ++k) {
    if (isnan(den[k])) { kmax = k; break; }
    if (den[k] > den[kmax]) kmax = k;
  }
  if (kmax == 0) return 2;
  for (int k = 0; k < kmax; ++k) if (den[k] < 0) return 1;
  double alt_min = alt[0];
  for (int k = 1; k < n_alt; ++k) if (alt[k] < alt_min) alt_min = alt[k];   /* lib:507 */

  double *m_own = NULL;
  const double *m = multiplier;
  if (!m) {
    m_own = (double *)malloc(sizeof(double) * (size_t)(n_points > 0 ? n_points : 1));
    vfo_oracle_multiplier(n_points, m_own);
    m = m_own;
  }
  row_out_t *rows = (row_out_t *)malloc(sizeof(row_out_t) * (size_t)(n_freq > 0 ? n_freq : 1));
  row_job_t job = {freq_mhz, alt, den, bmag, bpsi, kmax, mode, n_points, m, variant, rows, n_freq, 0};
  run_threads(row_worker, &job, n_threads, n_freq);
  /* lib:201: one decision for the whole call */
  double ymax = NAN;
  for (int r = 0; r < n_freq; ++r)
    if (!isnan(rows[r].ymax) && (isnan(ymax) || rows[r].ymax > ymax)) ymax = rows[r].ymax;
  int iso = (ymax < YTOL);                                /* NaN -> magnetised branch */
  for (int r = 0; r < n_freq; ++r) {
    double s = iso ? rows[r].s_iso : rows[r].s_mag;
    if (s == 0.0) s = NAN;                                 /* lib:290 */
    vh_out[r] = s + alt_min;                               /* lib:292 */
    if (hc_out) hc_out[r] = rows[r].hc;
  }
  free(rows);
  free(m_own);
  return 0;
}


typedef struct {
  const double *freq_mhz; int n_freq; int64_t freq_stride;
  const double *den, *bmag, *bpsi, *alt; int64_t alt_stride, n_profiles; int n_alt, mode, n_points;
  const double *multiplier; int variant; double *vh_out; int *status_out;
  int64_t next;
} batch_job_t;

static void *batch_worker(void *q) {
  batch_job_t *j = (batch_job_t *)q;
  for (;;) {
    int64_t p = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (p >= j->n_profiles) break;
    int st = vfo_oracle_profile(j->freq_mhz + p * j->freq_stride, j->n_freq, j->den + p * j->n_alt,
                                j->bmag + p * j->n_alt, j->bpsi + p * j->n_alt, j->alt + p * j->alt_stride,
                                j->n_alt, j->mode, j->n_points, j->multiplier, j->variant, 1,
                                j->vh_out + p * j->n_freq, NULL);
    if (st != 0)
      for (int r = 0; r < j->n_freq; ++r) j->vh_out[p * j->n_freq + r] = NAN;
    if (j->status_out) j->status_out[p] = st;
  }
  return NULL;
}

/*
 * Batch of profiles, den/bmag/bpsi [n_profiles x n_alt] row-major; alt and freq shared when
 * their stride is 0.  Profiles are distributed over OpenMP threads (rows inside one profile
 * run serially), which is how the CPU baseline uses every host core (n_threads <= 0: all).
 */
int vfo_oracle_batch(const double *freq_mhz, int n_freq, int64_t freq_stride, const double *den,
                     const double *bmag, const double *bpsi, const double *alt, int64_t alt_stride,
                     int64_t n_profiles, int n_alt, int mode, int n_points, const double *multiplier,
                     int variant, int n_threads, double *vh_out, int *status_out) {
  batch_job_t job = {freq_mhz, n_freq, freq_stride, den, bmag, bpsi, alt, alt_stride, n_profiles, n_alt,
                     mode, n_points, multiplier, variant, vh_out, status_out, 0};
  run_threads(batch_worker, &job, n_threads, n_profiles);
  return 0;
}

/* Elementwise helpers for known-answer tests (tests/test_core.py:137-152). */
void vfo_oracle_mup(const double *X, const double *Y, const double *psi, int n, int mode, int variant,
                    double *mup_out) {
  for (int i = 0; i < n; ++i)
    mup_out[i] = (variant == 0) ? mup_literal(X[i], Y[i], psi[i], mode == 0 ? 1.0 : -1.0)
                                : (double)mup_truth(X[i], Y[i], psi[i], mode);
}

double vfo_oracle_pairwise_sum(const double *a, long n) { return np_pairwise(a, n); }

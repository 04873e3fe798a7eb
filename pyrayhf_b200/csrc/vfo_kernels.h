// Internal launcher interface between the C ABI (vfo_cabi.cu) and the kernels (vfo_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace prhf {

constexpr int kThreads = 256;                // K1 block size
constexpr int kMaxWarps = 8;
#ifndef PRHF_TILE_THREADS
#define PRHF_TILE_THREADS 256
#endif
#ifndef PRHF_TILE_MINB
#define PRHF_TILE_MINB 4
#endif
constexpr int kTileThreads = PRHF_TILE_THREADS;   // K2 block size (multiple of 64, <= 256)
constexpr int kTileMinBlocks = PRHF_TILE_MINB;    // K2 resident CTAs per SM the register budget targets (64 registers:
                                                  // the grid loop itself does not spill, profiles/sweep_variants_r01c.log)
constexpr int kSoloMinBlocks = 3;                 // the single-launch kernel carries the row setup as well: 80 registers
constexpr int kRowsPerCta = kThreads / 32;   // K1: one warp per sounding frequency
constexpr int kRowWarpMaxPoints = 4096;      // direct-mode calls with n_points up to this use the row-per-warp kernel
// extra multiplier-table entries (value 1, weight 0) past n_points: the grid loop reads its tables one iteration
// (2 * kTileThreads points) ahead without a bounds test
constexpr int kMultPad = 2 * PRHF_TILE_THREADS + 4;
// entries of the multiplier table m (even, so that the weight table dm that follows it stays 16-byte aligned)
__host__ __device__ inline size_t mult_table_len(int n_points) { return ((size_t)n_points + kMultPad + 1) & ~(size_t)1; }

constexpr int kFlagIso = 1;       // unmagnetised branch (lib:201)
constexpr int kFlagGeneral = 2;   // non-finite node values / non-increasing altitudes / large angle steps
constexpr int kFlagFailed = 4;    // status != 0
constexpr int kFlagPsiConst = 8;  // field angle identical at every level below the peak
constexpr int kFlagPsiSmall = 16; // per-level field-angle steps <= kSmallRotateStep
constexpr int kFlagUniformAlt = 32; // levels within a quarter step of alt0 + k * mean step: the bracket guess is exact +-1
constexpr int kFlagAltUnsorted = 64; // altitudes below the peak not strictly increasing: np.interp's range tests decide

struct __align__(16) ProfileRecord {  // 64 bytes per profile, written by K1, read by K2
  int nt;                         // truncated length = argmax(den) (lib:371)
  int flags;
  double alt_min;                 // np.min(alt) (lib:507)
  double inv_dalt;                // (nt-1)/(alt[nt-1]-alt[0]); bracket guess for near-uniform grids
  double alt0;                    // alt[0]
  double sn0, cs0;                // sin / cos of the field angle at level 0 (all levels when kFlagPsiConst)
  double pad[2];
};

struct __align__(16) LiveRow {    // planned mode: one entry per row that reflects, appended by K1
  int row;                        // row index inside the launch
  int pad;
  double span;                    // h_c - alt0
};
// planner cost model (SM cycles): per-tile prologue + reduction, and loop cycles per grid point of one
// tile when the SM is fully occupied (measured with the developer phase trace, tools/trace_tiles.py)
constexpr float kPlanTilesPerSlot = 1.35f;        // planned mode: live tiles aimed at, per resident CTA slot
constexpr int kPlanMinTilePoints = 5000;          // ... and the shortest tile worth its prologue
constexpr int kMaxPlanCand = 32;

struct VfoParams {
  const double* freq;      // MHz; [n_freq] or per-profile rows
  int64_t freq_stride;     // 0 = shared
  int n_freq;
  const double* den;       // [P x n_alt]
  const double* bmag;
  const double* bpsi;
  const double* alt;       // [n_alt] or [P x n_alt]
  int64_t alt_stride;      // 0 = shared
  int n_alt;
  int64_t profile_offset;  // first profile handled by this launch
  const double* mult;      // stretched-grid multiplier m [mult_table_len(n_points)]
  const double* dmult;     // left-Riemann weights dm_i = m_{i+1} - m_i (0 from the last point on), same length
  const double* etab;      // E_i = exp(10 (1 - u_i)) (m_i = A - B E_i), same length: seeds of the E-space grid loop
  double e_ratio;          // exp(-10 * 2 kTileThreads / (n_points - 1)): E_{i + 2 kTileThreads} / E_i
  double e_weight;         // B (1 - exp(-10 / (n_points - 1))): dm_i = e_weight * E_i
  int n_points;
  int seg_len;             // grid points per tile (even)
  int n_seg;               // tiles per (profile, frequency) row
  int rows_per_warp;       // K1: sounding frequencies handled by one warp (CTA = 8 warps)
  int k1_solo;             // row setup inside the solo kernel: one row per CTA, scanned by warp 0
  int k1_lane_mode;        // K1: one thread per sounding frequency (large batches) instead of one warp
  int k1_finish_clamped;   // queued mode: K1 finishes rows clamped to the first level itself (1; 2 = literal arithmetic)
  int queue_cap_nodes;     // queued mode: levels of node buffer per CTA of the narrow kernel (live_count[1] = tickets,
                           // live_count[2] = rows deferred to the full-width kernel because their window is larger)
  LiveRow* defer_list;     // [rows_in_launch]
  int rw_rows_per_cta;     // row-per-warp kernel: rows of one profile handled by one CTA
  double* vh;              // [P x n_freq]
  int* status;             // [P] or null
  ProfileRecord* prof_rec; // [profiles_in_launch]
  double* row_span;        // [rows_in_launch]  h_c - alt0, NaN = row finished by K1
  double* partial;         // [rows_in_launch x n_seg] when n_seg > 1
  unsigned* counter;       // [rows_in_launch], zero on entry, zero on exit
  long long* trace;        // developer phase trace [tiles x 8] (PRHF_TRACE builds), else null
  // planned mode (small batches); null in direct mode
  unsigned* live_count;        // zeroed by a memset node before K1; K1 appends, K2 reads
  LiveRow* live_list;          // [rows_in_launch]
  int64_t rows_in_launch;
  int use_pdl;                 // launch the tile kernel with programmatic stream serialization
  int max_seg;             // stride of `partial` per row; planner's upper bound on n_seg
  int slots;               // resident tile-kernel CTAs on the device
  int n_sm, ctas_per_sm;   // its factors
  long long* trace_k1;     // developer phase trace of K1 [ctas x 8] (PRHF_TRACE builds), else null
  double freq_scale;       // row setup: f_hz = freq * freq_scale (1e6 for MHz input, lib:491; 1 for the Hz input of
                           // the standalone regrid stage).  The tile kernels always assume MHz.
  double* row_hc;          // optional [rows_in_launch]: reflection height h_c (lib:407), NaN on rows without one
  // profiles with more levels than the shared-memory staging holds (n_alt > prhf_max_n_alt): the row setup reads the
  // levels straight from global memory and the tile kernel reads un-scaled nodes from this per-profile table
  int levels_in_global;
  void* node_table;        // Node[profiles_in_launch x n_alt]
  // Kernel parameters sit in a constant bank that is cold at every launch; each 128-byte line of this struct costs
  // its first reader a miss.  Everything the single-profile path touches stays above this comment (two lines);
  // the planner's candidate tables (256 bytes, planned mode only) come last.
  int n_cand;              // planner candidates: (segments per row, grid points per segment)
  int cand_seg[kMaxPlanCand];
  int cand_len[kMaxPlanCand];
};

size_t vfo_smem_bytes(int n_alt);
int vfo_tile_ctas_per_sm(int n_alt, int max_smem_per_sm, bool solo_kernel);
cudaError_t launch_vfo_rows(const VfoParams& p, int mode, int64_t n_profiles, cudaStream_t stream);
cudaError_t launch_vfo_tiles(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream);
cudaError_t launch_vfo_rowwarp(const VfoParams& p, int mode, bool literal, int64_t n_ctas, cudaStream_t stream);
cudaError_t launch_vfo_solo(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream);
cudaError_t launch_vfo_queue(const VfoParams& p, int mode, bool literal, int64_t n_ctas, cudaStream_t stream);
int vfo_queue_ctas_per_sm();
int vfo_queue_threads();
int vfo_queue_cap_nodes(int n_alt, int max_smem_per_sm);
size_t vfo_node_bytes();
cudaError_t launch_vfo_nodes_global(const VfoParams& p, bool literal, int64_t n_profiles, cudaStream_t stream);
cudaError_t launch_vfo_tiles_global(const VfoParams& p, int mode, bool literal, int64_t n_tiles, cudaStream_t stream);
cudaError_t launch_grid_multiplier(int n, size_t n_padded, double* m, double* dm, double* e, cudaStream_t stream);
cudaError_t launch_mu_mup(const double* X, const double* Y, const double* psi, int64_t n, int mode, bool iso,
                          bool literal, double* mu, double* mup, cudaStream_t stream);
cudaError_t launch_residual(const double* vh, const double* vh_obs, int64_t n_profiles, int n_freq, double* residual,
                            double* chi2, cudaStream_t stream);
cudaError_t launch_argmin(const double* v, int64_t n, double* out2, cudaStream_t stream);
// standalone stages (vfo_stages.cu)
cudaError_t launch_den2freq(const double* den, int64_t n, double* out, int* negative_flag, cudaStream_t stream);
cudaError_t launch_find_x(const double* den, int64_t den_stride, const double* f_hz, int64_t f_stride, int64_t n,
                          double* X, int* negative_flag, cudaStream_t stream);
cudaError_t launch_find_y(const double* f_hz, int64_t f_stride, const double* b, int64_t b_stride, int64_t n, double* Y,
                          cudaStream_t stream);
cudaError_t launch_smooth_grid(double start, double end, int n_points, double sharpness, double* x, cudaStream_t stream);
struct RegridParams {
  const double* f_hz;      // [n_freq]
  int n_freq;
  const double *den, *bmag, *bpsi, *alt;   // one profile, [n_alt]
  const ProfileRecord* rec;                // from the row-setup kernel (truncation index)
  const double* row_hc;    // [n_freq]
  const double* mult;      // [n_points + kMultPad]
  int n_points;
  double *alt_out, *dist_out, *den_out, *bmag_out, *bpsi_out;   // [n_freq x n_points], any may be null
};
cudaError_t launch_regrid_write(const RegridParams& p, int n_alt, cudaStream_t stream);
cudaError_t launch_find_vh(const double* X, const double* Y, const double* psi, const double* dh, int64_t n_rows,
                           int64_t n_cols, double alt_min, int mode, bool literal, unsigned long long* scratch_word,
                           double* vh, cudaStream_t stream);
// stratified Snell's-law tracers (vfo_snell.cu)
struct SnellParams {
  const double* f0_hz;     // [n_rays]
  const double* elev_deg;  // [n_rays]
  int64_t n_rays;
  const double *alt, *ne, *babs, *bpsi;   // one profile, [n_alt]
  int n_alt;
  int mode, spherical, literal;
  double dz_target, apex_boost;
  int max_substeps;
  double r_e;
  double* scalars;         // [n_rays x 5]: group path km, group delay s, x midpoint, z midpoint, ground range km
  double* x_out;           // optional [n_rays x path_stride] (NaN-padded), with z_out
  double* z_out;
  int path_stride;
  int* n_path;             // optional [n_rays]: points on the path, 0 = no ray (every output NaN)
  // fan entry (prhf_snell_fan_f64): f0_hz [n_freq], elev_deg [rays_per_freq], ray = f * rays_per_freq + e, and the
  // refractive-index field computed once per frequency: field [n_freq x 2 x (n_alt + 1)] (mu, then mu').  0 / null: one
  // (f0, elevation) pair per ray and the field computed by every ray itself.
  int rays_per_freq;
  double* field;
};
size_t snell_smem_bytes(int n_alt);
cudaError_t launch_snell(const SnellParams& p, int max_smem_optin, cudaStream_t stream);
cudaError_t launch_synth_profiles(const double* params, int64_t n_profiles, const double* alt, int n_alt, double* den,
                                  double* bmag, double* bpsi, cudaStream_t stream);
cudaError_t launch_dfma_probe(double* out, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_math_selftest(int n, double* err2, cudaStream_t stream);

}  // namespace prhf

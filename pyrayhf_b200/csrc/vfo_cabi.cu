// C ABI of pyrayhf_b200 (declared in include/pyrayhf_b200.h).
//
// Replaces the reference's Python entry point PyRayHF.library.vertical_forward_operator
// (PyRayHF/library.py:459-509).  Host-side responsibilities kept here: argument validation with the
// reference's error behaviour (library.py:396 bad mode), caching of the stretched-grid multiplier
// table per n_points (library.py:361-364 recomputes it on every call), tile sizing, and the
// pinned-memory staging of the host-buffer entry point.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <new>
#include <vector>
#include <chrono>

#include "../../include/pyrayhf_b200.h"
#include "vfo_kernels.h"

struct GraphEntry {                  // host entry: captured [H2D copy, K1, K2] per call shape
  uint64_t epoch = 0;
  int calls = 0;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};   // one per planned-mode call parity
  int n_launches = 0;                              // kernels inside one replay
};

struct prhf_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  std::mutex mu;
  std::map<int, double*> mult;       // n_points -> device table
  // workspace for split rows (partials + self-resetting counters)
  double* partial = nullptr;
  unsigned* counter = nullptr;
  size_t partial_cap = 0, counter_cap = 0;
  // K1 -> K2 hand-off: one ProfileRecord per profile, one double per (profile, frequency) row
  prhf::ProfileRecord* prof_rec = nullptr;
  double* row_span = nullptr;
  size_t prof_cap = 0, row_cap = 0;
  // host entry: device arena + pinned mirror + private stream
  cudaStream_t stream = nullptr;
  char* d_arena = nullptr;
  char* h_arena = nullptr;
  size_t arena_cap = 0;
  int last_cuda_error = 0;
  int64_t launches = 0;
  int seg_len_override = 0;          // PRHF_SEG_LEN (tuning / tests)
  int64_t target_tiles = 0;          // PRHF_TARGET_TILES
  long long* trace = nullptr;        // developer phase trace buffer (PRHF_TRACE builds)
  size_t trace_k1_off = 0;           // K1 entries start here (in long longs)
  // planned mode (small batches): tile plan, compact tile list, K1 completion counter
  unsigned* live_count = nullptr;    // [0] live-row counter of planned mode,
                                     // [4..5] 8-byte scratch word of prhf_find_vh_f64
  bool use_solo = true;              // PRHF_NO_SOLO=1: single-profile calls through the two-kernel planned mode
  bool use_pdl = true;               // PRHF_NO_PDL=1: plain stream order between K1 and K2
  bool use_rowwarp = true;           // PRHF_NO_ROWWARP=1: small n_points through the tile kernel
  int force_nseg = 0;                // PRHF_PLAN_NSEG: planned mode uses exactly this many segments per row
  bool use_k1_lanes = true;          // PRHF_NO_K1_LANES=1: row setup always one warp per frequency
  int queue_mode = 1;                // PRHF_QUEUE=0: large batches with one tile-kernel CTA per row (no live-row queue)
  bool host_trace = false;                // PRHF_HOST_TRACE=1: wall-clock split of prhf_vfo_host_f64 on stderr
  double host_trace_us[4] = {0, 0, 0, 0};
  long host_trace_calls = 0;
  double* snell_field = nullptr;          // per-frequency refractive-index field of prhf_snell_fan_f64
  size_t snell_field_cap = 0;
  prhf::LiveRow* live_list = nullptr;     // [2 x live_list_cap]: the queue, then the rows deferred to the full-width kernel
  size_t live_list_cap = 0;
  void* node_table = nullptr;        // un-scaled nodes of profiles too long for shared memory (n_alt > prhf_max_n_alt)
  size_t node_table_cap = 0;
  // optional per-kernel timing (bench roofline): events around K1 and K2 of every launch pair
  bool kernel_timing = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  double k1_ms = 0.0, k2_ms = 0.0;
  int timed_pairs = 0;
  // host entry graph cache; `epoch` changes whenever a device buffer baked into a graph is reallocated
  uint64_t epoch = 1;
  bool use_graphs = true;            // PRHF_NO_GRAPH=1 disables
  std::map<std::array<int64_t, 9>, GraphEntry> graphs;
  int max_smem_per_sm = 0;
  int planned_max_rows = 4096;       // PRHF_PLANNED_MAX_ROWS
  // cross-stream guard: the K1 -> K2 hand-off buffers above are shared by every call on this ctx, so a call that
  // arrives on another stream than the previous one first waits (on the device) for that one to drain
  cudaEvent_t busy_ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool have_last_stream = false;
  // streaming entry (prhf_vfo_stream_f64): two device staging slots, a copy-in and a copy-out stream, events
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  cudaEvent_t ev_start = nullptr;
  char* slot_buf[2] = {nullptr, nullptr};
  size_t slot_cap = 0;
  char* shared_buf = nullptr;        // device copies of a shared freq / alt vector given on the host
  size_t shared_cap = 0;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

int fail(prhf_ctx* ctx, cudaError_t e) {
  if (ctx) ctx->last_cuda_error = (int)e;
  cudaGetLastError();  // clear sticky-less errors
  return PRHF_ERR_CUDA;
}

#define PRHF_CUDA(ctx, call)                         \
  do {                                               \
    cudaError_t _e = (call);                         \
    if (_e != cudaSuccess) return fail((ctx), _e);   \
  } while (0)

// Constants of the E-space grid loop (tile_sum_fast_e): the third table and the two ratios of the geometric sequence
// E_i = exp(10 (1 - i / (n - 1))).
void set_stretch_constants(prhf::VfoParams& P, const double* mult, int n_points) {
  P.etab = mult + 2 * prhf::mult_table_len(n_points);
  const double step = (n_points > 1) ? 10.0 / (double)(n_points - 1) : 0.0;
  P.e_ratio = std::exp(-step * (double)(2 * prhf::kTileThreads));
  P.e_weight = -std::expm1(-step) / (std::exp(10.0) - 1.0);
}

int get_multiplier(prhf_ctx* ctx, int n_points, cudaStream_t stream, const double** out) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto it = ctx->mult.find(n_points);
  if (it != ctx->mult.end()) {
    *out = it->second;
    return PRHF_OK;
  }
  double* m = nullptr;
  const size_t len = prhf::mult_table_len(n_points);          // [m | dm | E], dm = m + len, E = m + 2 len
  PRHF_CUDA(ctx, cudaMalloc(&m, sizeof(double) * 3 * len));
  cudaError_t e = prhf::launch_grid_multiplier(n_points, len, m, m + len, m + 2 * len, stream);
  ctx->launches += 2;
  if (e != cudaSuccess) {
    cudaFree(m);
    return fail(ctx, e);
  }
  // Make the table visible to every later stream: the cache outlives this call.
  PRHF_CUDA(ctx, cudaStreamSynchronize(stream));
  ctx->mult[n_points] = m;
  *out = m;
  return PRHF_OK;
}

int ensure_records(prhf_ctx* ctx, size_t n_prof, size_t n_rows) {
  if (n_prof > ctx->prof_cap) {
    if (ctx->prof_rec) cudaFree(ctx->prof_rec);
    ctx->prof_rec = nullptr;
    ctx->prof_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->prof_rec, sizeof(prhf::ProfileRecord) * n_prof));
    ctx->prof_cap = n_prof;
    ctx->epoch++;
  }
  if (n_rows > ctx->row_cap) {
    if (ctx->row_span) cudaFree(ctx->row_span);
    ctx->row_span = nullptr;
    ctx->row_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->row_span, sizeof(double) * n_rows));
    ctx->row_cap = n_rows;
    ctx->epoch++;
  }
  return PRHF_OK;
}

int ensure_plan(prhf_ctx* ctx, size_t n_rows) {
  if (!ctx->live_count) {
    PRHF_CUDA(ctx, cudaMalloc(&ctx->live_count, 8 * sizeof(unsigned)));
    PRHF_CUDA(ctx, cudaMemset(ctx->live_count, 0, 8 * sizeof(unsigned)));
  }
  if (n_rows > ctx->live_list_cap) {
    if (ctx->live_list) cudaFree(ctx->live_list);
  if (ctx->snell_field) cudaFree(ctx->snell_field);
    ctx->live_list = nullptr;
    ctx->live_list_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->live_list, sizeof(prhf::LiveRow) * 2 * n_rows));
    ctx->live_list_cap = n_rows;
    ctx->epoch++;
  }
  return PRHF_OK;
}

int ensure_workspace(prhf_ctx* ctx, size_t n_partial, size_t n_counter) {
  if (n_partial > ctx->partial_cap) {
    if (ctx->partial) cudaFree(ctx->partial);
    ctx->partial = nullptr;
    ctx->partial_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->partial, sizeof(double) * n_partial));
    ctx->partial_cap = n_partial;
    ctx->epoch++;
  }
  if (n_counter > ctx->counter_cap) {
    // A larger counter array replaces the old one; in-flight launches on the old array must finish.
    PRHF_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->counter) cudaFree(ctx->counter);
    ctx->counter = nullptr;
    ctx->counter_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->counter, sizeof(unsigned) * n_counter));
    PRHF_CUDA(ctx, cudaMemset(ctx->counter, 0, sizeof(unsigned) * n_counter));
    ctx->counter_cap = n_counter;
    ctx->epoch++;
  }
  return PRHF_OK;
}

// Tile sizing: split each (profile, frequency) row into n_seg segments so that a launch has enough
// CTAs to fill 148 SMs several times over, but never below 1024 grid points per tile (per-tile setup
// -- profile staging and the critical-curve scan -- is amortised over the tile's points).
void choose_tiling(const prhf_ctx* ctx, int64_t rows, int n_points, int* seg_len, int* n_seg) {
  if (ctx->seg_len_override > 0) {
    int sl = std::min(ctx->seg_len_override, std::max(n_points, 1));
    sl += (sl & 1);          // tiles start on even grid indices (two points per thread per iteration)
    *seg_len = sl;
    *n_seg = (n_points + sl - 1) / sl;
    return;
  }
  const int64_t target = ctx->target_tiles > 0 ? ctx->target_tiles : (int64_t)ctx->sm_count * 16;
  int64_t want = (target + rows - 1) / std::max<int64_t>(rows, 1);
  const int max_seg = std::max(1, n_points / 1024);
  want = std::max<int64_t>(1, std::min<int64_t>(want, max_seg));
  int sl = (int)((n_points + want - 1) / want);
  sl = ((sl + 2 * prhf::kTileThreads - 1) / (2 * prhf::kTileThreads)) * (2 * prhf::kTileThreads);
  sl = std::max(sl, 1);
  *seg_len = sl;
  *n_seg = (n_points + sl - 1) / sl;
}

bool use_planned_mode(const prhf_ctx* ctx, int64_t rows_total, int n_points) {
  return ctx->seg_len_override <= 0 && rows_total <= (int64_t)ctx->planned_max_rows && n_points >= 2048;
}

// How a call is decomposed: solo (single launch, one profile), planned (small batch), else direct.
struct CallMode {
  bool solo, planned;
  int slots, ctas_per_sm;
};
CallMode call_mode(const prhf_ctx* ctx, int64_t rows_total, int n_points, int n_alt) {
  CallMode m;
  const bool small = use_planned_mode(ctx, rows_total, n_points);
  // the single-launch kernel keeps fewer CTAs resident (it carries the row setup): decide with its own slot count
  const int solo_ctas = prhf::vfo_tile_ctas_per_sm(n_alt, ctx->max_smem_per_sm, true);
  m.solo = ctx->use_solo && small && rows_total * 2 <= (int64_t)ctx->sm_count * solo_ctas;
  m.ctas_per_sm = m.solo ? solo_ctas : prhf::vfo_tile_ctas_per_sm(n_alt, ctx->max_smem_per_sm, false);
  m.slots = ctx->sm_count * m.ctas_per_sm;
  m.planned = small && !m.solo;
  return m;
}

int validate(const prhf_ctx* ctx, const void* freq, int n_freq, const void* den, const void* bmag, const void* bpsi,
             const void* alt, int64_t n_profiles, int n_alt, int mode, int n_points, const void* vh) {
  if (!ctx) return PRHF_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return PRHF_ERR_BAD_MODE;
  if (n_freq < 0 || n_profiles < 0 || n_alt < 1 || n_points < 1) return PRHF_ERR_INVALID_ARG;
  if (n_freq == 0 || n_profiles == 0) return PRHF_OK;
  if (!freq || !den || !bmag || !bpsi || !alt || !vh) return PRHF_ERR_INVALID_ARG;
  return PRHF_OK;
}

// More levels than the shared-memory staging holds: the operator switches to its global-memory form (no limit, as
// np.interp has none, lib:424-426); the regrid stage and the tracers still stage in shared memory.
bool too_long_for_smem(const prhf_ctx* ctx, int n_alt) {
  return prhf::vfo_smem_bytes(n_alt) > (size_t)ctx->max_smem_optin;
}

int ensure_node_table(prhf_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->node_table_cap) return PRHF_OK;
  PRHF_CUDA(ctx, cudaDeviceSynchronize());
  if (ctx->node_table) cudaFree(ctx->node_table);
  ctx->node_table = nullptr;
  ctx->node_table_cap = 0;
  PRHF_CUDA(ctx, cudaMalloc(&ctx->node_table, bytes));
  ctx->node_table_cap = bytes;
  ctx->epoch++;
  return PRHF_OK;
}

}  // namespace

extern "C" {

int prhf_version(void) { return PRHF_VERSION; }

const char* prhf_error_string(int code) {
  switch (code) {
    case PRHF_OK: return "ok";
    case PRHF_ERR_INVALID_ARG: return "invalid argument";
    case PRHF_ERR_BAD_MODE: return "mode must be 'O' or 'X'";
    case PRHF_ERR_CUDA: return "CUDA runtime error";
    case PRHF_ERR_NALT_TOO_LARGE: return "n_alt exceeds the shared-memory staging limit";
    case PRHF_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
    default: return "unknown error";
  }
}

int prhf_last_cuda_error(const prhf_ctx* ctx, const char** text) {
  const int e = ctx ? ctx->last_cuda_error : 0;
  if (text) *text = cudaGetErrorString((cudaError_t)e);
  return e;
}

int prhf_ctx_create(int device, prhf_ctx** out) {
  if (!out) return PRHF_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return PRHF_ERR_NO_DEVICE;
  }
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return PRHF_ERR_NO_DEVICE;
  if (device >= count) return PRHF_ERR_INVALID_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PRHF_ERR_NO_DEVICE;
  if (prop.major != 10) return PRHF_ERR_NO_DEVICE;   // the only code object in this library is sm_100a
  prhf_ctx* ctx = new (std::nothrow) prhf_ctx();
  if (!ctx) return PRHF_ERR_INVALID_ARG;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->max_smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
  if (const char* s = getenv("PRHF_PLANNED_MAX_ROWS")) ctx->planned_max_rows = atoi(s);
  if (const char* s = getenv("PRHF_NO_GRAPH")) ctx->use_graphs = (atoi(s) == 0);
  if (const char* s = getenv("PRHF_NO_PDL")) ctx->use_pdl = (atoi(s) == 0);
  if (const char* s = getenv("PRHF_NO_SOLO")) ctx->use_solo = (atoi(s) == 0);
  if (const char* s = getenv("PRHF_NO_ROWWARP")) ctx->use_rowwarp = (atoi(s) == 0);
  if (const char* s = getenv("PRHF_NO_K1_LANES")) ctx->use_k1_lanes = (atoi(s) == 0);
  if (const char* s = getenv("PRHF_QUEUE")) ctx->queue_mode = atoi(s);
  if (const char* s = getenv("PRHF_HOST_TRACE")) ctx->host_trace = (atoi(s) != 0);
  if (const char* s = getenv("PRHF_PLAN_NSEG")) ctx->force_nseg = atoi(s);
  if (const char* s = getenv("PRHF_SEG_LEN")) ctx->seg_len_override = atoi(s);
  if (const char* s = getenv("PRHF_TARGET_TILES")) ctx->target_tiles = atoll(s);
  DeviceGuard g(device);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return PRHF_ERR_CUDA;
  }
  *out = ctx;
  return PRHF_OK;
}

void prhf_ctx_destroy(prhf_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  for (int k = 0; k < 3; ++k)
    if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
  for (auto& kv : ctx->graphs)
    for (int k = 0; k < 2; ++k)
      if (kv.second.exec[k]) cudaGraphExecDestroy(kv.second.exec[k]);
  for (auto& kv : ctx->mult) cudaFree(kv.second);
  if (ctx->partial) cudaFree(ctx->partial);
  if (ctx->counter) cudaFree(ctx->counter);
  if (ctx->live_count) cudaFree(ctx->live_count);
  if (ctx->live_list) cudaFree(ctx->live_list);
  if (ctx->node_table) cudaFree(ctx->node_table);
  if (ctx->prof_rec) cudaFree(ctx->prof_rec);
  if (ctx->row_span) cudaFree(ctx->row_span);
  if (ctx->d_arena) cudaFree(ctx->d_arena);
  if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
  for (int k = 0; k < 2; ++k) {
    if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
    if (ctx->ev_k[k]) cudaEventDestroy(ctx->ev_k[k]);
    if (ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
    if (ctx->slot_buf[k]) cudaFree(ctx->slot_buf[k]);
  }
  if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
  if (ctx->busy_ev) cudaEventDestroy(ctx->busy_ev);
  if (ctx->shared_buf) cudaFree(ctx->shared_buf);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int prhf_max_n_alt(const prhf_ctx* ctx) {
  if (!ctx) return 0;
  return (int)((size_t)ctx->max_smem_optin / prhf::vfo_smem_bytes(1));
}

int64_t prhf_launch_count(const prhf_ctx* ctx) { return ctx ? ctx->launches : 0; }

int prhf_kernel_timing(prhf_ctx* ctx, int enable, double* rows_kernel_ms, double* tile_kernel_ms, int* launch_pairs) {
  if (!ctx) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  if (rows_kernel_ms) *rows_kernel_ms = ctx->k1_ms;
  if (tile_kernel_ms) *tile_kernel_ms = ctx->k2_ms;
  if (launch_pairs) *launch_pairs = ctx->timed_pairs;
  ctx->k1_ms = ctx->k2_ms = 0.0;
  ctx->timed_pairs = 0;
  if (enable && !ctx->ev[0])
    for (int k = 0; k < 3; ++k) PRHF_CUDA(ctx, cudaEventCreate(&ctx->ev[k]));
  ctx->kernel_timing = enable != 0;
  return PRHF_OK;
}

#ifdef PRHF_TRACE
// developer-only: allocate / read back the tile-kernel phase trace (not declared in the public header)
int prhf_debug_trace_alloc(prhf_ctx* ctx, int64_t n_tiles) {
  DeviceGuard g(ctx->device);
  if (ctx->trace) cudaFree(ctx->trace);
  ctx->trace = nullptr;
  PRHF_CUDA(ctx, cudaMalloc(&ctx->trace, sizeof(long long) * 8 * (size_t)(n_tiles + 4096)));
  PRHF_CUDA(ctx, cudaMemset(ctx->trace, 0, sizeof(long long) * 8 * (size_t)(n_tiles + 4096)));
  ctx->trace_k1_off = 8 * (size_t)n_tiles;
  return PRHF_OK;
}
int prhf_debug_trace_read(prhf_ctx* ctx, int64_t n_tiles, long long* out) {
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, cudaDeviceSynchronize());
  PRHF_CUDA(ctx, cudaMemcpy(out, ctx->trace, sizeof(long long) * 8 * (size_t)(n_tiles + 4096), cudaMemcpyDeviceToHost));
  return PRHF_OK;
}
#endif

int prhf_grid_multiplier_f64(prhf_ctx* ctx, int n_points, double* m_out, void* cuda_stream) {
  if (!ctx || n_points < 1 || !m_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_grid_multiplier(n_points, (size_t)n_points, m_out, nullptr, nullptr, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

// The launch sequence of one call (row setup + tiles, or the single-launch kernel) on `cuda_stream`.  Shared by the
// device entry, the host entry (under stream capture) and the streaming entry.
static int vfo_enqueue(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride, const double* den,
                       const double* bmag, const double* bpsi, const double* alt, int64_t alt_profile_stride,
                       int64_t n_profiles, int n_alt, int mode, int n_points, unsigned flags, double* vh_out, int* status,
                       void* cuda_stream) {
  int rc = validate(ctx, freq_mhz, n_freq, den, bmag, bpsi, alt, n_profiles, n_alt, mode, n_points, vh_out);
  if (rc != PRHF_OK) return rc;
  if (n_freq == 0 || n_profiles == 0) return PRHF_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const double* mult = nullptr;
  rc = get_multiplier(ctx, n_points, stream, &mult);
  if (rc != PRHF_OK) return rc;

  const int64_t rows_total = n_profiles * (int64_t)n_freq;
  const bool literal = (flags & PRHF_FLAG_LITERAL) != 0;
  const bool big = too_long_for_smem(ctx, n_alt);             // global-memory form: direct mode only
  CallMode cm = call_mode(ctx, rows_total, n_points, big ? 1 : n_alt);
  if (big) {
    cm.solo = cm.planned = false;
    cm.ctas_per_sm = prhf::kTileMinBlocks;
    cm.slots = ctx->sm_count * cm.ctas_per_sm;
  }
  const int ctas_per_sm = cm.ctas_per_sm, slots = cm.slots;

  // Small batches (fewer rows than a few waves of tiles): planned mode.  K1's last CTA counts the rows
  // that reflect and sizes the segments so that the live tiles fill the resident-CTA slots; K2 strides
  // over the compact tile list.  Large batches: direct mode, one tile per row.
  // Single profile (rows * 2 <= resident CTAs): solo mode, one launch with a static split of every row.
  const bool solo = cm.solo, planned = cm.planned;
  int seg_len = 0, n_seg = 0;
  int n_cand = 0, cand_seg[prhf::kMaxPlanCand] = {0}, cand_len[prhf::kMaxPlanCand] = {0};
  if (solo) {
    const int quantum = 2 * prhf::kTileThreads;
    int want = (int)std::min<int64_t>(slots / rows_total, std::max(1, n_points / 2048));
    for (;; --want) {                                         // largest split whose segments are all non-empty
      seg_len = ((n_points + want - 1) / want + quantum - 1) / quantum * quantum;
      n_seg = (n_points + seg_len - 1) / seg_len;
      if (n_seg == want || want <= 1) break;
    }
  } else if (planned) {
    n_seg = std::max(1, std::min(prhf::kMaxPlanCand, n_points / 1024));   // upper bound (stride of the partials)
    seg_len = n_points;                                       // unused by the kernels in planned mode
    // candidate tilings: ns segments of a length that is a multiple of two points per thread
    const int quantum = 2 * prhf::kTileThreads;
    for (int ns = 1; ns <= n_seg && n_cand < prhf::kMaxPlanCand; ++ns) {
      if (ctx->force_nseg > 0 && ns != ctx->force_nseg) continue;   // developer override (PRHF_PLAN_NSEG)
      const int sl = ((n_points + ns - 1) / ns + quantum - 1) / quantum * quantum;
      if ((n_points + sl - 1) / sl != ns) continue;           // rounding made a segment empty
      cand_seg[n_cand] = ns;
      cand_len[n_cand] = sl;
      ++n_cand;
    }
  } else {
    choose_tiling(ctx, rows_total, n_points, &seg_len, &n_seg);
  }

  // A launch covers at most max_tiles tiles (grid.x limit) and max_rows rows (hand-off workspace);
  // profiles are chunked accordingly.
  const int64_t max_tiles = (int64_t)1 << 30;
  const int64_t max_rows = (int64_t)1 << 24;
  const int64_t tiles_per_profile = (int64_t)n_freq * n_seg;
  if (tiles_per_profile > max_tiles) return PRHF_ERR_INVALID_ARG;
  int64_t prof_per_launch = std::max<int64_t>(1, std::min(max_tiles / tiles_per_profile, max_rows / n_freq));
  if (big)                                                    // node table of one launch: at most 512 MiB
    prof_per_launch = std::max<int64_t>(1, std::min<int64_t>(prof_per_launch,
                                                             (int64_t)(((size_t)512 << 20) / (prhf::vfo_node_bytes() * (size_t)n_alt))));
  prof_per_launch = std::min(prof_per_launch, n_profiles);
  const int64_t rows_launch = prof_per_launch * n_freq;
  rc = ensure_records(ctx, (size_t)prof_per_launch, (size_t)rows_launch);
  if (rc != PRHF_OK) return rc;
  if (n_seg > 1) {
    rc = ensure_workspace(ctx, (size_t)rows_launch * n_seg, (size_t)rows_launch);
    if (rc != PRHF_OK) return rc;
  }
  // Queued mode (large batches through the tile kernel): the row setup, one thread per frequency, finishes the rows
  // clamped to the first level itself and appends the rows that still need their grid points to a queue; the tile
  // kernel runs one CTA per resident slot and strides over the queue.  Against one CTA per row this drops the launch of
  // a CTA for every row that does not reflect (two of three on a global grid) and keeps every queued row equally long.
  // (not in solo mode: the single-launch kernel scans its one row with the whole block -- 148 ... 222 profiles of ONE
  //  frequency used to take the thread-per-frequency branch there and left their reflecting rows unwritten)
  const bool lane_k1 = !planned && !solo && ctx->use_k1_lanes && n_profiles >= (int64_t)ctx->sm_count;
  const bool queued = !solo && !planned && !big && lane_k1 && ctx->queue_mode > 0 && n_seg == 1 &&
                      !(n_points <= prhf::kRowWarpMaxPoints && ctx->use_rowwarp);
  if (queued) {
    n_cand = 1;
    cand_seg[0] = n_seg;
    cand_len[0] = seg_len;
  }
  if (planned || queued) {
    rc = ensure_plan(ctx, (size_t)rows_launch);
    if (rc != PRHF_OK) return rc;
  }
  if (big) {
    rc = ensure_node_table(ctx, prhf::vfo_node_bytes() * (size_t)n_alt * (size_t)prof_per_launch);
    if (rc != PRHF_OK) return rc;
  }
  // K1 granularity: 8 rows per CTA while the launch is small (latency matters), whole profiles per CTA
  // once there are enough profiles to fill the GPU (amortises the per-CTA profile staging).
  const int chunks8 = (n_freq + prhf::kRowsPerCta - 1) / prhf::kRowsPerCta;
  int rows_per_warp = 1;
  if (n_profiles * chunks8 > (int64_t)ctx->sm_count * 64) {
    rows_per_warp = (int)std::min<int64_t>(chunks8, (n_profiles * chunks8) / ((int64_t)ctx->sm_count * 32));
    rows_per_warp = std::max(rows_per_warp, 1);
  }
  for (int64_t p0 = 0; p0 < n_profiles; p0 += prof_per_launch) {
    const int64_t np = std::min(prof_per_launch, n_profiles - p0);
    prhf::VfoParams P;
    P.freq = freq_mhz;
    P.freq_stride = freq_profile_stride;
    P.n_freq = n_freq;
    P.den = den;
    P.bmag = bmag;
    P.bpsi = bpsi;
    P.alt = alt;
    P.alt_stride = alt_profile_stride;
    P.n_alt = n_alt;
    P.profile_offset = p0;
    P.mult = mult;
    P.dmult = mult + prhf::mult_table_len(n_points);
    set_stretch_constants(P, mult, n_points);
    P.n_points = n_points;
    P.seg_len = seg_len;
    P.n_seg = n_seg;
    P.rows_per_warp = solo ? 1 : rows_per_warp;
    P.k1_solo = solo ? 1 : 0;
    P.k1_lane_mode = lane_k1 ? 1 : 0;
    P.k1_finish_clamped = queued ? (literal ? 2 : 1) : 0;
    P.queue_cap_nodes = queued ? prhf::vfo_queue_cap_nodes(n_alt, ctx->max_smem_per_sm) : 0;
    P.defer_list = queued ? ctx->live_list + ctx->live_list_cap : nullptr;
    P.vh = vh_out;
    P.status = status;
    P.prof_rec = ctx->prof_rec;
    P.row_span = ctx->row_span;
    P.partial = ctx->partial;
    P.counter = ctx->counter;
    P.trace = ctx->trace;
    if (planned || queued) {
      // the live-row counter starts every launch at zero (a 4-byte memset node: simpler and safer than any
      // hand-over of the reset between consecutive calls)
      PRHF_CUDA(ctx, cudaMemsetAsync(ctx->live_count, 0, 4 * sizeof(unsigned), stream));   // [live rows, tickets, deferred rows]
      P.live_count = ctx->live_count;
      P.live_list = ctx->live_list;
    } else {
      P.live_count = nullptr;
      P.live_list = nullptr;
    }
    P.rows_in_launch = np * n_freq;
    // Programmatic dependent launch lets the tile grid start while the row-setup grid drains.  In planned mode that
    // early start places the working CTAs on the SMs the row setup has already left and leaves the others short of
    // tiles (profiles/trace of 4 profiles: 271 tiles on 64 SMs), which costs more than the overlap gains as soon as
    // the row setup spans more than a couple of profiles (measured, 2 ... 23 profiles: better without from 3 on, up to
    // 12 % at 12 profiles).
    P.use_pdl = (ctx->use_pdl && !(planned && rows_total > 2 * (int64_t)n_freq)) ? 1 : 0;
    P.max_seg = n_seg;
    P.slots = slots;
    P.n_sm = ctx->sm_count;
    P.ctas_per_sm = ctas_per_sm;
    P.n_cand = n_cand;
    for (int c = 0; c < prhf::kMaxPlanCand; ++c) { P.cand_seg[c] = cand_seg[c]; P.cand_len[c] = cand_len[c]; }
    P.trace_k1 = ctx->trace ? ctx->trace + ctx->trace_k1_off : nullptr;
    P.freq_scale = 1e6;                                       // lib:491
    P.row_hc = nullptr;
    P.levels_in_global = big ? 1 : 0;
    P.node_table = ctx->node_table;
    if (big) {
      // row setup (thread per frequency, levels read in place) -> node table -> tiles; plain stream order
      P.use_pdl = 0;
      P.k1_lane_mode = 1;
      const bool timing = ctx->kernel_timing;
      if (timing) PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[0], stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_rows(P, mode, np, stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_nodes_global(P, literal, np, stream));
      if (timing) PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[1], stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_tiles_global(P, mode, literal, np * tiles_per_profile, stream));
      ctx->launches += 3;
      if (timing) {
        PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[2], stream));
        PRHF_CUDA(ctx, cudaEventSynchronize(ctx->ev[2]));
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
        ctx->k1_ms += a;
        ctx->k2_ms += b;
        ctx->timed_pairs++;
      }
      continue;
    }
    if (solo) {
      if (ctx->kernel_timing) PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[1], stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_solo(P, mode, literal, np * tiles_per_profile, stream));
      ctx->launches += 1;
      if (ctx->kernel_timing) {                               // the one kernel counts as the "tile kernel"
        PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[2], stream));
        PRHF_CUDA(ctx, cudaEventSynchronize(ctx->ev[2]));
        float b = 0.f;
        cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
        ctx->k2_ms += b;
        ctx->timed_pairs++;
      }
      continue;
    }
    // direct mode with few grid points per row: row-per-warp kernel (profile staged once per CTA)
    const bool rowwarp = !planned && n_seg == 1 && n_points <= prhf::kRowWarpMaxPoints && ctx->use_rowwarp;
    int64_t rw_ctas = 0;
    if (rowwarp) {
      // whole profiles per CTA when there are enough profiles, else split the rows to get >= 4 waves of CTAs
      int64_t chunks = std::max<int64_t>(1, std::min<int64_t>(chunks8, ((int64_t)4 * slots + np - 1) / np));
      P.rw_rows_per_cta = (int)((n_freq + chunks - 1) / chunks);
      chunks = (n_freq + P.rw_rows_per_cta - 1) / P.rw_rows_per_cta;
      rw_ctas = np * chunks;
    } else {
      P.rw_rows_per_cta = n_freq;
    }
    // planned mode: enough CTAs for two waves of slots; they stride over however many tiles K1 planned
    const int64_t grid = planned ? std::min<int64_t>(np * tiles_per_profile, (int64_t)2 * slots) : np * tiles_per_profile;
    // queued mode: the narrow kernel (one CTA per slot, eight slots per SM, tickets) over the queue, then the full-width
    // kernel over the rows the narrow one deferred (window larger than its node buffer; usually none: the launch is a
    // few microseconds of CTAs that read a zero and leave)
    auto launch_queued = [&]() -> int {
      prhf::VfoParams Q = P;
      const double step = (n_points > 1) ? 10.0 / (double)(n_points - 1) : 0.0;
      Q.e_ratio = std::exp(-step * (double)(2 * prhf::vfo_queue_threads()));
      const int64_t q_slots = (int64_t)ctx->sm_count * prhf::vfo_queue_ctas_per_sm();
      PRHF_CUDA(ctx, prhf::launch_vfo_queue(Q, mode, literal, std::min<int64_t>(np * (int64_t)n_freq, q_slots), stream));
      prhf::VfoParams D = P;
      D.live_count = ctx->live_count + 2;
      D.live_list = P.defer_list;
      D.use_pdl = 0;
      PRHF_CUDA(ctx, prhf::launch_vfo_tiles(D, mode, literal, std::min<int64_t>(np * (int64_t)n_freq, (int64_t)slots), stream));
      ctx->launches += 1;                                     // (the row-setup + first tile launch are counted by the caller)
      return PRHF_OK;
    };
    if (rowwarp && !ctx->kernel_timing) {
      PRHF_CUDA(ctx, prhf::launch_vfo_rows(P, mode, np, stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_rowwarp(P, mode, literal, rw_ctas, stream));
      ctx->launches += 2;
      continue;
    }
    if (ctx->kernel_timing) {
      // measurement mode: events between the kernels (this serialises them: no PDL overlap)
      P.use_pdl = 0;
      PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[0], stream));
      PRHF_CUDA(ctx, prhf::launch_vfo_rows(P, mode, np, stream));
      PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[1], stream));
      if (rowwarp) PRHF_CUDA(ctx, prhf::launch_vfo_rowwarp(P, mode, literal, rw_ctas, stream));
      else if (queued) { rc = launch_queued(); if (rc != PRHF_OK) return rc; }
      else PRHF_CUDA(ctx, prhf::launch_vfo_tiles(P, mode, literal, grid, stream));
      PRHF_CUDA(ctx, cudaEventRecord(ctx->ev[2], stream));
      PRHF_CUDA(ctx, cudaEventSynchronize(ctx->ev[2]));
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
      cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
      ctx->k1_ms += a;
      ctx->k2_ms += b;
      ctx->timed_pairs++;
      ctx->launches += 2;
      continue;
    }
    PRHF_CUDA(ctx, prhf::launch_vfo_rows(P, mode, np, stream));
    if (queued) { rc = launch_queued(); if (rc != PRHF_OK) return rc; }
    else PRHF_CUDA(ctx, prhf::launch_vfo_tiles(P, mode, literal, grid, stream));
    ctx->launches += 2;
  }
  return PRHF_OK;
}

// The hand-off buffers of a ctx (ProfileRecord, row spans, partials, counters, live list) serve one launch sequence
// at a time.  Calls that arrive on different streams are therefore ordered on the device: the newcomer's stream
// waits for the event recorded behind the previous call.  Same-stream calls pay one event record (~1 us of host time).
static int stream_guard_enter(prhf_ctx* ctx, cudaStream_t stream) {
  if (ctx->have_last_stream && ctx->last_stream != stream && ctx->busy_ev)
    PRHF_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->busy_ev, 0));
  return PRHF_OK;
}
static int stream_guard_leave(prhf_ctx* ctx, cudaStream_t stream) {
  if (!ctx->busy_ev) PRHF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->busy_ev, cudaEventDisableTiming));
  PRHF_CUDA(ctx, cudaEventRecord(ctx->busy_ev, stream));
  ctx->last_stream = stream;
  ctx->have_last_stream = true;
  return PRHF_OK;
}

int prhf_vfo_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride, const double* den,
                 const double* bmag, const double* bpsi, const double* alt, int64_t alt_profile_stride,
                 int64_t n_profiles, int n_alt, int mode, int n_points, unsigned flags, double* vh_out, int* status,
                 void* cuda_stream) {
  int rc = validate(ctx, freq_mhz, n_freq, den, bmag, bpsi, alt, n_profiles, n_alt, mode, n_points, vh_out);
  if (rc != PRHF_OK) return rc;
  if (n_freq == 0 || n_profiles == 0) return PRHF_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  rc = stream_guard_enter(ctx, stream);
  if (rc != PRHF_OK) return rc;
  rc = vfo_enqueue(ctx, freq_mhz, n_freq, freq_profile_stride, den, bmag, bpsi, alt, alt_profile_stride, n_profiles, n_alt,
                   mode, n_points, flags, vh_out, status, cuda_stream);
  if (rc != PRHF_OK) return rc;
  return stream_guard_leave(ctx, stream);
}

int prhf_vfo_host_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride,
                      const double* den, const double* bmag, const double* bpsi, const double* alt,
                      int64_t alt_profile_stride, int64_t n_profiles, int n_alt, int mode, int n_points,
                      unsigned flags, double* vh_out, int* status) {
  int rc = validate(ctx, freq_mhz, n_freq, den, bmag, bpsi, alt, n_profiles, n_alt, mode, n_points, vh_out);
  if (rc != PRHF_OK) return rc;
  if (n_freq == 0 || n_profiles == 0) return PRHF_OK;
  DeviceGuard g(ctx->device);
  rc = stream_guard_enter(ctx, ctx->stream);                  // a device-entry call may still be in flight elsewhere
  if (rc != PRHF_OK) return rc;

  // Profiles are processed in chunks through one packed arena:
  //   inputs  [freq | alt | den | bmag | bpsi]   (one H2D copy per chunk)
  //   outputs [vh | status]                      (one D2H copy per chunk)
  const size_t d8 = sizeof(double);
  const bool freq_shared = (freq_profile_stride == 0), alt_shared = (alt_profile_stride == 0);
  const size_t per_prof_in = d8 * ((freq_shared ? 0 : (size_t)n_freq) + (alt_shared ? 0 : (size_t)n_alt) + 3 * (size_t)n_alt);
  const size_t per_prof_out = d8 * (size_t)n_freq + sizeof(int);
  const size_t shared_in = d8 * ((freq_shared ? (size_t)n_freq : 0) + (alt_shared ? (size_t)n_alt : 0));
  const size_t budget = (size_t)256 << 20;
  int64_t chunk = (int64_t)std::max<size_t>(1, (budget - std::min(budget / 2, shared_in)) / (per_prof_in + per_prof_out));
  chunk = std::min<int64_t>(chunk, n_profiles);
  const size_t in_bytes = shared_in + per_prof_in * (size_t)chunk;
  const size_t out_off = (in_bytes + 255) & ~(size_t)255;
  const size_t vh_bytes = d8 * (size_t)n_freq * (size_t)chunk;
  const size_t need = out_off + vh_bytes + sizeof(int) * (size_t)chunk + 256;
  if (need > ctx->arena_cap) {
    PRHF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
    ctx->d_arena = ctx->h_arena = nullptr;
    ctx->arena_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->d_arena, need));
    PRHF_CUDA(ctx, cudaMallocHost(&ctx->h_arena, need));
    ctx->arena_cap = need;
    ctx->epoch++;
  }
  // Small calls (one chunk, outputs <= 1 MiB): the kernels write vh / status straight into the pinned host
  // arena (zero-copy over PCIe, no D2H copy node) and the [H2D copy, K1, K2] sequence is replayed from a
  // CUDA graph captured on the second call with the same shape.
  const bool small = (chunk == n_profiles) && (vh_bytes + sizeof(int) * (size_t)chunk <= ((size_t)1 << 20));
  for (int64_t p0 = 0; p0 < n_profiles; p0 += chunk) {
    const int64_t np = std::min(chunk, n_profiles - p0);
    size_t off = 0;
    // copy `rows` rows of `count` doubles (source row stride `stride` doubles) densely into the arena
    auto put = [&](const double* src, int64_t stride, size_t count, int64_t rows) {
      const size_t at = off;
      if (rows == 1 || stride == (int64_t)count) {
        memcpy(ctx->h_arena + off, src, d8 * count * (size_t)rows);
      } else {
        for (int64_t q = 0; q < rows; ++q) memcpy(ctx->h_arena + off + d8 * count * (size_t)q, src + q * stride, d8 * count);
      }
      off += d8 * count * (size_t)rows;
      return at;
    };
    const bool host_trace = ctx->host_trace;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t o_freq = freq_shared ? put(freq_mhz, 0, (size_t)n_freq, 1)
                                      : put(freq_mhz + p0 * freq_profile_stride, freq_profile_stride, (size_t)n_freq, np);
    const size_t o_alt = alt_shared ? put(alt, 0, (size_t)n_alt, 1)
                                    : put(alt + p0 * alt_profile_stride, alt_profile_stride, (size_t)n_alt, np);
    const size_t o_den = put(den + p0 * n_alt, n_alt, (size_t)n_alt, np);
    const size_t o_b = put(bmag + p0 * n_alt, n_alt, (size_t)n_alt, np);
    const size_t o_psi = put(bpsi + p0 * n_alt, n_alt, (size_t)n_alt, np);
    const auto t1 = std::chrono::steady_clock::now();
    char* out_base = small ? ctx->h_arena : ctx->d_arena;       // pinned host memory is device-addressable (UVA)
    double* d_vh = (double*)(out_base + out_off);
    int* d_st = (int*)(out_base + out_off + d8 * (size_t)n_freq * np);
    auto enqueue = [&]() -> int {
      PRHF_CUDA(ctx, cudaMemcpyAsync(ctx->d_arena, ctx->h_arena, off, cudaMemcpyHostToDevice, ctx->stream));
      return vfo_enqueue(ctx, (const double*)(ctx->d_arena + o_freq), n_freq, freq_shared ? 0 : n_freq,
                          (const double*)(ctx->d_arena + o_den), (const double*)(ctx->d_arena + o_b),
                          (const double*)(ctx->d_arena + o_psi), (const double*)(ctx->d_arena + o_alt),
                          alt_shared ? 0 : n_alt, np, n_alt, mode, n_points, flags, d_vh, d_st, ctx->stream);
    };
    bool done = false;
    if (small && ctx->use_graphs) {
      const std::array<int64_t, 9> key = {n_freq, freq_shared, alt_shared, np, n_alt, mode, n_points, (int64_t)flags,
                                          ctx->seg_len_override};
      if (ctx->graphs.size() >= 64 && ctx->graphs.find(key) == ctx->graphs.end()) {
        // a caller cycling through many shapes: drop the cache rather than let it grow without bound
        for (auto& kv : ctx->graphs)
          for (int k = 0; k < 2; ++k)
            if (kv.second.exec[k]) cudaGraphExecDestroy(kv.second.exec[k]);
        ctx->graphs.clear();
      }
      GraphEntry& ge = ctx->graphs[key];
      if (ge.epoch != ctx->epoch) {                           // buffers moved since capture: start over
        for (int k = 0; k < 2; ++k)
          if (ge.exec[k]) { cudaGraphExecDestroy(ge.exec[k]); ge.exec[k] = nullptr; }
        ge.calls = 0;
      }
      // (one executable per call shape: alternating two of them does not shorten cudaGraphLaunch in back-to-back calls,
      //  profiles/host_trace_r02.txt)
      const int parity = 0;
      if (ge.calls >= 1) {
        if (!ge.exec[parity]) {
          cudaGraph_t graph = nullptr;
          if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const int64_t launches0 = ctx->launches;
            rc = enqueue();
            ge.n_launches = (int)(ctx->launches - launches0);
            ctx->launches = launches0;
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc == PRHF_OK && ce == cudaSuccess && graph &&
                cudaGraphInstantiate(&ge.exec[parity], graph, 0) != cudaSuccess)
              ge.exec[parity] = nullptr;
            if (graph) cudaGraphDestroy(graph);
          }
          if (!ge.exec[parity]) {                              // capture is an optimisation, never a requirement
            cudaGetLastError();
            ctx->use_graphs = false;
          }
        }
        if (ge.exec[parity]) {
          PRHF_CUDA(ctx, cudaGraphLaunch(ge.exec[parity], ctx->stream));
          ctx->launches += ge.n_launches;
          done = true;
        }
      }
      if (!done) {
        rc = enqueue();
        if (rc != PRHF_OK) return rc;
        ge.calls++;
        ge.epoch = ctx->epoch;
        done = true;
      }
    }
    if (!done) {
      rc = enqueue();
      if (rc != PRHF_OK) return rc;
    }
    const size_t out_bytes = d8 * (size_t)n_freq * np + sizeof(int) * (size_t)np;
    if (!small)
      PRHF_CUDA(ctx, cudaMemcpyAsync(ctx->h_arena + out_off, ctx->d_arena + out_off, out_bytes, cudaMemcpyDeviceToHost,
                                     ctx->stream));
    const auto t2 = std::chrono::steady_clock::now();
    PRHF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const auto t3 = std::chrono::steady_clock::now();
    memcpy(vh_out + p0 * n_freq, ctx->h_arena + out_off, d8 * (size_t)n_freq * np);
    if (status) memcpy(status + p0, ctx->h_arena + out_off + d8 * (size_t)n_freq * np, sizeof(int) * (size_t)np);
    if (host_trace) {                                         // PRHF_HOST_TRACE=1: where a host-entry call spends its time
      const auto t4 = std::chrono::steady_clock::now();
      auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
      };
      ctx->host_trace_us[0] += us(t0, t1);
      ctx->host_trace_us[1] += us(t1, t2);
      ctx->host_trace_us[2] += us(t2, t3);
      ctx->host_trace_us[3] += us(t3, t4);
      if (++ctx->host_trace_calls % 64 == 0) {
        const double n = 64.0;
        fprintf(stderr, "prhf host trace (mean of 64 calls, us): pack %.2f  enqueue %.2f  wait %.2f  unpack %.2f\n",
                ctx->host_trace_us[0] / n, ctx->host_trace_us[1] / n, ctx->host_trace_us[2] / n, ctx->host_trace_us[3] / n);
        for (double& v : ctx->host_trace_us) v = 0.0;
      }
    }
  }
  ctx->have_last_stream = false;                              // everything on this ctx has drained
  return PRHF_OK;
}

// ---- streaming entry -------------------------------------------------------------------------------------------
namespace {
bool on_device(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
// rows x count doubles, source row stride src_stride, destination row stride dst_stride (both in doubles)
cudaError_t copy_rows(double* dst, int64_t dst_stride, const double* src, int64_t src_stride, size_t count, int64_t rows,
                      cudaMemcpyKind kind, cudaStream_t st) {
  if (rows <= 0 || count == 0) return cudaSuccess;
  if ((src_stride == (int64_t)count && dst_stride == (int64_t)count) || rows == 1)
    return cudaMemcpyAsync(dst, src, sizeof(double) * count * (size_t)rows, kind, st);
  return cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)dst_stride, src, sizeof(double) * (size_t)src_stride,
                           sizeof(double) * count, (size_t)rows, kind, st);
}
}  // namespace

int prhf_vfo_stream_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride,
                        const double* den, const double* bmag, const double* bpsi, const double* alt,
                        int64_t alt_profile_stride, int64_t n_profiles, int n_alt, int mode, int n_points,
                        unsigned flags, int64_t chunk_profiles, double* vh_out, int64_t vh_profile_stride, int* status,
                        void* cuda_stream, int synchronize) {
  int rc = validate(ctx, freq_mhz, n_freq, den, bmag, bpsi, alt, n_profiles, n_alt, mode, n_points, vh_out);
  if (rc != PRHF_OK) return rc;
  if (n_freq == 0 || n_profiles == 0) return PRHF_OK;
  if (vh_profile_stride == 0) vh_profile_stride = n_freq;
  if (vh_profile_stride < n_freq || chunk_profiles < 0) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  cudaStream_t sk = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  const size_t d8 = sizeof(double);
  const bool freq_shared = (freq_profile_stride == 0), alt_shared = (alt_profile_stride == 0);
  const bool den_dev = on_device(den), b_dev = on_device(bmag), psi_dev = on_device(bpsi);
  const bool freq_dev = on_device(freq_mhz), alt_dev = on_device(alt);
  const bool vh_dev = on_device(vh_out), st_dev = status ? on_device(status) : true;
  const bool vh_direct = vh_dev && vh_profile_stride == n_freq;   // the kernels write the caller's buffer themselves
  rc = stream_guard_enter(ctx, sk);
  if (rc != PRHF_OK) return rc;

  // chunking: enough chunks for the copies to hide behind the kernels, chunks large enough to fill the GPU
  int64_t chunk = chunk_profiles;
  const size_t per_prof = d8 * ((den_dev ? 0 : (size_t)n_alt) + (b_dev ? 0 : (size_t)n_alt) + (psi_dev ? 0 : (size_t)n_alt) +
                                ((alt_shared || alt_dev) ? 0 : (size_t)n_alt) +
                                ((freq_shared || freq_dev) ? 0 : (size_t)n_freq) + (vh_direct ? 0 : (size_t)n_freq)) +
                          ((status && st_dev) ? 0 : sizeof(int));
  if (chunk == 0) {
    // (every chunk ends with the tail of a persistent tile kernel, ~40 us of draining SMs, and a launch gap: at most
    //  about five chunks per call once the opening quarter-chunk hides the first copy-in)
    chunk = std::min<int64_t>(4096, std::max<int64_t>(256, (n_profiles + 3) / 4));
    if (per_prof == 0) chunk = n_profiles;                    // nothing to stage: one launch sequence
  }
  if (per_prof > 0) chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, (int64_t)(((size_t)512 << 20) / per_prof)));
  chunk = std::min<int64_t>(chunk, n_profiles);
  // Chunk boundaries.  The copy-in of the FIRST chunk is the one copy nothing can hide, so (unless the caller fixed the
  // chunk size) the call opens with a quarter-size chunk: 0.3 ms instead of 1.1 ms of exposed PCIe time per call at
  // 4096 profiles per chunk, and the kernels of that short chunk already cover the copy-in of the next, full one.
  std::vector<int64_t> starts;
  {
    int64_t p = 0;
    if (chunk_profiles == 0 && per_prof > 0 && n_profiles > chunk && chunk >= 1024) {
      starts.push_back(0);
      p = chunk / 4;
    }
    for (; p < n_profiles; p += chunk) starts.push_back(p);
    starts.push_back(n_profiles);
  }
  const int64_t n_chunks = (int64_t)starts.size() - 1;

  // resources: streams, events, two staging slots, device copies of shared host vectors
  if (!ctx->s_in) {
    PRHF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    PRHF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      PRHF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
      PRHF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_k[k], cudaEventDisableTiming));
      PRHF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[k], cudaEventDisableTiming));
    }
    PRHF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
  }
  const size_t slot_bytes = ((per_prof * (size_t)chunk + 255) & ~(size_t)255) + 8 * 256;   // carve() rounds to 256 B
  if (slot_bytes > ctx->slot_cap) {
    PRHF_CUDA(ctx, cudaDeviceSynchronize());
    for (int k = 0; k < 2; ++k) {
      if (ctx->slot_buf[k]) cudaFree(ctx->slot_buf[k]);
      ctx->slot_buf[k] = nullptr;
    }
    ctx->slot_cap = 0;
    for (int k = 0; k < 2; ++k) PRHF_CUDA(ctx, cudaMalloc(&ctx->slot_buf[k], slot_bytes));
    ctx->slot_cap = slot_bytes;
  }
  const size_t shared_bytes = d8 * (((freq_shared && !freq_dev) ? (size_t)n_freq : 0) + ((alt_shared && !alt_dev) ? (size_t)n_alt : 0));
  if (shared_bytes > ctx->shared_cap) {
    PRHF_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->shared_buf) cudaFree(ctx->shared_buf);
    ctx->shared_buf = nullptr;
    ctx->shared_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->shared_buf, shared_bytes + 256));
    ctx->shared_cap = shared_bytes;
  }

  // everything below is ordered after what the caller already enqueued on the compute stream (device inputs may
  // still be in production there)
  PRHF_CUDA(ctx, cudaEventRecord(ctx->ev_start, sk));
  PRHF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
  PRHF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
  const double* d_freq_shared = freq_mhz;
  const double* d_alt_shared = alt;
  {
    char* sb = ctx->shared_buf;
    if (freq_shared && !freq_dev) {
      PRHF_CUDA(ctx, cudaMemcpyAsync(sb, freq_mhz, d8 * (size_t)n_freq, cudaMemcpyHostToDevice, ctx->s_in));
      d_freq_shared = (const double*)sb;
      sb += d8 * (size_t)n_freq;
    }
    if (alt_shared && !alt_dev) {
      PRHF_CUDA(ctx, cudaMemcpyAsync(sb, alt, d8 * (size_t)n_alt, cudaMemcpyHostToDevice, ctx->s_in));
      d_alt_shared = (const double*)sb;
    }
  }

  // Developer timeline (PRHF_STREAM_TRACE=1): timing events around the three phases of every chunk, printed to
  // stderr after the call (forces synchronisation).
  const bool trace = getenv("PRHF_STREAM_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  mark(sk);                                                   // t = 0

  struct Staged {
    const double *den, *b, *psi, *alt, *freq;
    int64_t alt_stride, freq_stride;
    double* vh;
    int* st;
  } staged[2];

  // Issue order is breadth-first: the copy-in of chunk c+1 is enqueued BEFORE the kernels of chunk c, so the copy can
  // never sit behind them in a hardware queue two streams happen to share.
  auto issue_copy_in = [&](int64_t c) -> int {
    const int slot = (int)(c & 1);
    const int64_t p0 = starts[c], np = starts[c + 1] - p0;
    char* base = ctx->slot_buf[slot];
    size_t off = 0;
    auto carve = [&](size_t bytes) {
      char* q = base + off;
      off += (bytes + 255) & ~(size_t)255;
      return q;
    };
    // the slot's inputs are free once the kernels of chunk c-2 have finished
    if (c >= 2) PRHF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_k[slot], 0));
    mark(ctx->s_in);
    auto stage = [&](const double* src, bool dev, int64_t stride, size_t count, const double** out) -> cudaError_t {
      if (dev) {
        *out = src + p0 * stride;
        return cudaSuccess;
      }
      double* dst = (double*)carve(d8 * count * (size_t)np);
      *out = dst;
      return copy_rows(dst, (int64_t)count, src + p0 * stride, stride, count, np, cudaMemcpyHostToDevice, ctx->s_in);
    };
    Staged& S = staged[slot];
    S.alt = d_alt_shared;
    S.freq = d_freq_shared;
    S.alt_stride = S.freq_stride = 0;
    PRHF_CUDA(ctx, stage(den, den_dev, n_alt, (size_t)n_alt, &S.den));
    PRHF_CUDA(ctx, stage(bmag, b_dev, n_alt, (size_t)n_alt, &S.b));
    PRHF_CUDA(ctx, stage(bpsi, psi_dev, n_alt, (size_t)n_alt, &S.psi));
    if (!alt_shared) {
      PRHF_CUDA(ctx, stage(alt, alt_dev, alt_profile_stride, (size_t)n_alt, &S.alt));
      S.alt_stride = alt_dev ? alt_profile_stride : n_alt;
    }
    if (!freq_shared) {
      PRHF_CUDA(ctx, stage(freq_mhz, freq_dev, freq_profile_stride, (size_t)n_freq, &S.freq));
      S.freq_stride = freq_dev ? freq_profile_stride : n_freq;
    }
    mark(ctx->s_in);
    PRHF_CUDA(ctx, cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
    S.vh = vh_direct ? vh_out + p0 * vh_profile_stride : (double*)carve(d8 * (size_t)n_freq * (size_t)np);
    S.st = status ? (st_dev ? status + p0 : (int*)carve(sizeof(int) * (size_t)np)) : nullptr;
    return PRHF_OK;
  };

  rc = issue_copy_in(0);
  if (rc != PRHF_OK) return rc;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int slot = (int)(c & 1);
    const int64_t p0 = starts[c], np = starts[c + 1] - p0;
    if (c + 1 < n_chunks) {
      rc = issue_copy_in(c + 1);
      if (rc != PRHF_OK) return rc;
    }
    const Staged& S = staged[slot];
    // ---- compute stream: inputs of this chunk have landed, outputs of chunk c-2 have left the slot ----
    PRHF_CUDA(ctx, cudaStreamWaitEvent(sk, ctx->ev_in[slot], 0));
    if (c >= 2) PRHF_CUDA(ctx, cudaStreamWaitEvent(sk, ctx->ev_out[slot], 0));
    mark(sk);
    rc = vfo_enqueue(ctx, S.freq, n_freq, S.freq_stride, S.den, S.b, S.psi, S.alt, S.alt_stride, np, n_alt, mode, n_points,
                     flags, S.vh, S.st, sk);
    if (rc != PRHF_OK) return rc;
    mark(sk);
    PRHF_CUDA(ctx, cudaEventRecord(ctx->ev_k[slot], sk));
    // ---- copy-out stream ----
    PRHF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[slot], 0));
    mark(ctx->s_out);
    if (!vh_direct)
      PRHF_CUDA(ctx, copy_rows(vh_out + p0 * vh_profile_stride, vh_profile_stride, S.vh, n_freq, (size_t)n_freq, np,
                               vh_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->s_out));
    if (status && !st_dev)
      PRHF_CUDA(ctx, cudaMemcpyAsync(status + p0, S.st, sizeof(int) * (size_t)np, cudaMemcpyDeviceToHost, ctx->s_out));
    mark(ctx->s_out);
    PRHF_CUDA(ctx, cudaEventRecord(ctx->ev_out[slot], ctx->s_out));
  }
  // the compute stream ends behind the last copies, so "after this call" in stream order means "results delivered"
  PRHF_CUDA(ctx, cudaStreamWaitEvent(sk, ctx->ev_out[(n_chunks - 1) & 1], 0));
  if (n_chunks >= 2) PRHF_CUDA(ctx, cudaStreamWaitEvent(sk, ctx->ev_out[(n_chunks - 2) & 1], 0));
  rc = stream_guard_leave(ctx, sk);
  if (rc != PRHF_OK) return rc;
  if (synchronize || trace) {
    PRHF_CUDA(ctx, cudaStreamSynchronize(sk));
    ctx->have_last_stream = false;
  }
  if (trace) {
    // event order: [0] start, then per copy-in 2 marks, per chunk 2 compute + 2 copy-out marks, in issue order
    fprintf(stderr, "prhf stream trace: %lld chunks of %lld profiles (ms since the call started)\n", (long long)n_chunks,
            (long long)chunk);
    std::vector<float> t(tev.size(), 0.f);
    for (size_t k = 1; k < tev.size(); ++k) cudaEventElapsedTime(&t[k], tev[0], tev[k]);
    // issue order: in(0): 1,2 ; then for c: [in(c+1): 2 marks if any], k(c): 2, out(c): 2
    size_t k = 1;
    std::vector<float> in_b(n_chunks), in_e(n_chunks), k_b(n_chunks), k_e(n_chunks), o_b(n_chunks), o_e(n_chunks);
    in_b[0] = t[k++]; in_e[0] = t[k++];
    for (int64_t c = 0; c < n_chunks; ++c) {
      if (c + 1 < n_chunks) { in_b[c + 1] = t[k++]; in_e[c + 1] = t[k++]; }
      k_b[c] = t[k++]; k_e[c] = t[k++]; o_b[c] = t[k++]; o_e[c] = t[k++];
    }
    for (int64_t c = 0; c < n_chunks; ++c)
      fprintf(stderr, "  chunk %2lld  copy-in %8.3f -> %8.3f   kernels %8.3f -> %8.3f   copy-out %8.3f -> %8.3f\n",
              (long long)c, in_b[c], in_e[c], k_b[c], k_e[c], o_b[c], o_e[c]);
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  return PRHF_OK;
}

// Page-lock / release a host range the caller owns (the shared-memory result buffer of the sharded operator).
int prhf_host_register(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return PRHF_ERR_INVALID_ARG;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return PRHF_OK;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return PRHF_ERR_CUDA;
  }
  return PRHF_OK;
}
int prhf_host_unregister(void* ptr) {
  if (!ptr) return PRHF_ERR_INVALID_ARG;
  if (cudaHostUnregister(ptr) != cudaSuccess) {
    cudaGetLastError();
    return PRHF_ERR_CUDA;
  }
  return PRHF_OK;
}

int prhf_mu_mup_f64(prhf_ctx* ctx, const double* X, const double* Y, const double* bpsi_deg, int64_t n, int mode,
                    int isotropic, unsigned flags, double* mu_out, double* mup_out, void* cuda_stream) {
  if (!ctx || n < 0) return PRHF_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return PRHF_ERR_BAD_MODE;
  if (n == 0) return PRHF_OK;
  if (!X || (!isotropic && (!Y || !bpsi_deg))) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_mu_mup(X, Y, bpsi_deg, n, mode, isotropic != 0, (flags & PRHF_FLAG_LITERAL) != 0, mu_out,
                                     mup_out, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

// ---- standalone stages (vfo_stages.cu); device pointers, asynchronous on the stream ----
int prhf_den2freq_f64(prhf_ctx* ctx, const double* density, int64_t n, double* freq_out, int* negative_flag,
                      void* cuda_stream) {
  if (!ctx || n < 0) return PRHF_ERR_INVALID_ARG;
  if (n == 0) return PRHF_OK;
  if (!density || !freq_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_den2freq(density, n, freq_out, negative_flag, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_find_x_f64(prhf_ctx* ctx, const double* n_e, int64_t n_e_stride, const double* f_hz, int64_t f_stride,
                    int64_t n, double* x_out, int* negative_flag, void* cuda_stream) {
  if (!ctx || n < 0 || (n_e_stride != 0 && n_e_stride != 1) || (f_stride != 0 && f_stride != 1))
    return PRHF_ERR_INVALID_ARG;
  if (n == 0) return PRHF_OK;
  if (!n_e || !f_hz || !x_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_find_x(n_e, n_e_stride, f_hz, f_stride, n, x_out, negative_flag,
                                     (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_find_y_f64(prhf_ctx* ctx, const double* f_hz, int64_t f_stride, const double* b, int64_t b_stride, int64_t n,
                    double* y_out, void* cuda_stream) {
  if (!ctx || n < 0 || (b_stride != 0 && b_stride != 1) || (f_stride != 0 && f_stride != 1))
    return PRHF_ERR_INVALID_ARG;
  if (n == 0) return PRHF_OK;
  if (!f_hz || !b || !y_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_find_y(f_hz, f_stride, b, b_stride, n, y_out, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_smooth_grid_f64(prhf_ctx* ctx, double start, double end, int n_points, double sharpness, double* x_out,
                         void* cuda_stream) {
  if (!ctx || n_points < 0) return PRHF_ERR_INVALID_ARG;
  if (n_points == 0) return PRHF_OK;
  if (!x_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_smooth_grid(start, end, n_points, sharpness, x_out, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_regrid_f64(prhf_ctx* ctx, const double* f_hz, int n_freq, const double* n_e, const double* b,
                    const double* bpsi, const double* aalt, int n_alt, int mode, int n_points, double* crit_height,
                    double* alt_out, double* dist_out, double* den_out, double* bmag_out, double* bpsi_out, int* status,
                    void* cuda_stream) {
  int rc = validate(ctx, f_hz, n_freq, n_e, b, bpsi, aalt, 1, n_alt, mode, n_points, crit_height);
  if (rc != PRHF_OK) return rc;
  if (n_freq == 0) return PRHF_OK;
  if (too_long_for_smem(ctx, n_alt)) return PRHF_ERR_NALT_TOO_LARGE;   // this stage stages the profile in shared memory
  if (n_freq > 65535) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const double* mult = nullptr;
  rc = get_multiplier(ctx, n_points, stream, &mult);
  if (rc != PRHF_OK) return rc;
  rc = ensure_records(ctx, 1, (size_t)2 * n_freq);            // row spans + a scratch row for the setup kernel's vh
  if (rc != PRHF_OK) return rc;
  prhf::VfoParams P{};
  P.freq = f_hz;
  P.freq_scale = 1.0;                                         // this stage takes Hz (lib:333-334)
  P.n_freq = n_freq;
  P.den = n_e;
  P.bmag = b;
  P.bpsi = bpsi;
  P.alt = aalt;
  P.n_alt = n_alt;
  P.mult = mult;
  P.dmult = mult + prhf::mult_table_len(n_points);
  set_stretch_constants(P, mult, n_points);
  P.n_points = n_points;
  P.seg_len = n_points;
  P.n_seg = 1;
  P.rows_per_warp = 1;
  P.vh = ctx->row_span + n_freq;
  P.status = status;
  P.prof_rec = ctx->prof_rec;
  P.row_span = ctx->row_span;
  P.row_hc = crit_height;
  P.rows_in_launch = n_freq;
  P.max_seg = 1;
  P.n_sm = ctx->sm_count;
  PRHF_CUDA(ctx, prhf::launch_vfo_rows(P, mode, 1, stream));
  prhf::RegridParams R;
  R.f_hz = f_hz;
  R.n_freq = n_freq;
  R.den = n_e;
  R.bmag = b;
  R.bpsi = bpsi;
  R.alt = aalt;
  R.rec = ctx->prof_rec;
  R.row_hc = crit_height;
  R.mult = mult;
  R.n_points = n_points;
  R.alt_out = alt_out;
  R.dist_out = dist_out;
  R.den_out = den_out;
  R.bmag_out = bmag_out;
  R.bpsi_out = bpsi_out;
  ctx->launches++;
  if (alt_out || dist_out || den_out || bmag_out || bpsi_out) {
    PRHF_CUDA(ctx, prhf::launch_regrid_write(R, n_alt, stream));
    ctx->launches++;
  }
  return PRHF_OK;
}

int prhf_find_vh_f64(prhf_ctx* ctx, const double* X, const double* Y, const double* bpsi_deg, const double* dh,
                     int64_t n_rows, int64_t n_cols, double alt_min, int mode, unsigned flags, double* vh_out,
                     void* cuda_stream) {
  if (!ctx || n_rows < 0 || n_cols < 0) return PRHF_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return PRHF_ERR_BAD_MODE;
  if (n_rows == 0) return PRHF_OK;
  if (!vh_out || (n_cols > 0 && (!X || !Y || !bpsi_deg || !dh))) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  int rc = ensure_plan(ctx, 1);                               // allocates the scratch word
  if (rc != PRHF_OK) return rc;
  PRHF_CUDA(ctx, prhf::launch_find_vh(X, Y, bpsi_deg, dh, n_rows, n_cols, alt_min, mode,
                                      (flags & PRHF_FLAG_LITERAL) != 0,
                                      reinterpret_cast<unsigned long long*>(ctx->live_count + 4), vh_out,
                                      (cudaStream_t)cuda_stream));
  ctx->launches += 2;
  return PRHF_OK;
}

int prhf_synth_profiles_f64(prhf_ctx* ctx, const double* params, int64_t n_profiles, const double* alt, int n_alt,
                            double* den_out, double* bmag_out, double* bpsi_out, void* cuda_stream) {
  if (!ctx || n_profiles < 0 || n_alt < 0) return PRHF_ERR_INVALID_ARG;
  if (n_profiles == 0 || n_alt == 0) return PRHF_OK;
  if (!params || !alt || !den_out || !bmag_out || !bpsi_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_synth_profiles(params, n_profiles, alt, n_alt, den_out, bmag_out, bpsi_out,
                                             (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_snell_f64(prhf_ctx* ctx, const double* f0_hz, const double* elevation_deg, int64_t n_rays,
                   const double* alt_km, const double* ne, const double* babs, const double* bpsi, int n_alt, int mode,
                   int geometry, unsigned flags, double dz_target_km, double apex_boost, int max_substeps, double r_e_km,
                   double* scalars_out, double* x_out, double* z_out, int path_stride, int* n_path_out,
                   void* cuda_stream) {
  if (!ctx || n_rays < 0 || n_alt < 1) return PRHF_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return PRHF_ERR_BAD_MODE;
  if (geometry != 0 && geometry != 1) return PRHF_ERR_INVALID_ARG;
  if (n_rays == 0) return PRHF_OK;
  if (!f0_hz || !elevation_deg || !alt_km || !ne || !babs || !bpsi || !scalars_out) return PRHF_ERR_INVALID_ARG;
  if ((x_out == nullptr) != (z_out == nullptr)) return PRHF_ERR_INVALID_ARG;
  if (x_out && path_stride < 2 * (n_alt + 1) + 1) return PRHF_ERR_INVALID_ARG;
  if (geometry == 1 && (!(dz_target_km > 0.0) || max_substeps < 1)) return PRHF_ERR_INVALID_ARG;
  if (prhf::snell_smem_bytes(n_alt) > (size_t)ctx->max_smem_optin) return PRHF_ERR_NALT_TOO_LARGE;
  DeviceGuard g(ctx->device);
  prhf::SnellParams P;
  P.f0_hz = f0_hz;
  P.elev_deg = elevation_deg;
  P.n_rays = n_rays;
  P.alt = alt_km;
  P.ne = ne;
  P.babs = babs;
  P.bpsi = bpsi;
  P.n_alt = n_alt;
  P.mode = mode;
  P.spherical = geometry;
  P.literal = (flags & PRHF_FLAG_LITERAL) ? 1 : 0;
  P.dz_target = dz_target_km;
  P.apex_boost = apex_boost;
  P.max_substeps = max_substeps;
  P.r_e = r_e_km;
  P.scalars = scalars_out;
  P.x_out = x_out;
  P.z_out = z_out;
  P.path_stride = path_stride;
  P.n_path = n_path_out;
  P.rays_per_freq = 0;
  P.field = nullptr;
  PRHF_CUDA(ctx, prhf::launch_snell(P, ctx->max_smem_optin, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_snell_fan_f64(prhf_ctx* ctx, const double* f0_hz, int n_freq, const double* elevation_deg, int n_elev,
                       const double* alt_km, const double* ne, const double* babs, const double* bpsi, int n_alt,
                       int mode, int geometry, unsigned flags, double dz_target_km, double apex_boost, int max_substeps,
                       double r_e_km, double* scalars_out, double* x_out, double* z_out, int path_stride,
                       int* n_path_out, void* cuda_stream) {
  if (!ctx || n_freq < 0 || n_elev < 0 || n_alt < 1) return PRHF_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return PRHF_ERR_BAD_MODE;
  if (geometry != 0 && geometry != 1) return PRHF_ERR_INVALID_ARG;
  if (n_freq == 0 || n_elev == 0) return PRHF_OK;
  if (!f0_hz || !elevation_deg || !alt_km || !ne || !babs || !bpsi || !scalars_out) return PRHF_ERR_INVALID_ARG;
  if ((x_out == nullptr) != (z_out == nullptr)) return PRHF_ERR_INVALID_ARG;
  if (x_out && path_stride < 2 * (n_alt + 1) + 1) return PRHF_ERR_INVALID_ARG;
  if (geometry == 1 && (!(dz_target_km > 0.0) || max_substeps < 1)) return PRHF_ERR_INVALID_ARG;
  if (prhf::snell_smem_bytes(n_alt) > (size_t)ctx->max_smem_optin) return PRHF_ERR_NALT_TOO_LARGE;
  DeviceGuard g(ctx->device);
  // the per-frequency field table lives in the ctx (grown on demand; in-flight launches on the old one must finish)
  const size_t need = sizeof(double) * 2 * (size_t)(n_alt + 1) * (size_t)n_freq;
  if (need > ctx->snell_field_cap) {
    PRHF_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->snell_field) cudaFree(ctx->snell_field);
    ctx->snell_field = nullptr;
    ctx->snell_field_cap = 0;
    PRHF_CUDA(ctx, cudaMalloc(&ctx->snell_field, need));
    ctx->snell_field_cap = need;
  }
  prhf::SnellParams P;
  P.f0_hz = f0_hz;
  P.elev_deg = elevation_deg;
  P.n_rays = (int64_t)n_freq * n_elev;
  P.alt = alt_km;
  P.ne = ne;
  P.babs = babs;
  P.bpsi = bpsi;
  P.n_alt = n_alt;
  P.mode = mode;
  P.spherical = geometry;
  P.literal = (flags & PRHF_FLAG_LITERAL) ? 1 : 0;
  P.dz_target = dz_target_km;
  P.apex_boost = apex_boost;
  P.max_substeps = max_substeps;
  P.r_e = r_e_km;
  P.scalars = scalars_out;
  P.x_out = x_out;
  P.z_out = z_out;
  P.path_stride = path_stride;
  P.n_path = n_path_out;
  P.rays_per_freq = n_elev;
  P.field = ctx->snell_field;
  PRHF_CUDA(ctx, prhf::launch_snell(P, ctx->max_smem_optin, (cudaStream_t)cuda_stream));
  ctx->launches += 2;
  return PRHF_OK;
}

int prhf_residual_f64(prhf_ctx* ctx, const double* vh_model, const double* vh_obs, int64_t n_profiles, int n_freq,
                      double* residual_out, double* chi2_out, void* cuda_stream) {
  if (!ctx || n_profiles < 0 || n_freq < 0) return PRHF_ERR_INVALID_ARG;
  if (n_profiles == 0 || n_freq == 0) return PRHF_OK;
  if (!vh_model || !vh_obs || (!residual_out && !chi2_out)) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_residual(vh_model, vh_obs, n_profiles, n_freq, residual_out, chi2_out,
                                       (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_argmin_f64(prhf_ctx* ctx, const double* values, int64_t n, double* out2, void* cuda_stream) {
  if (!ctx || n < 0 || !out2 || (n > 0 && !values)) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  PRHF_CUDA(ctx, prhf::launch_argmin(values, n, out2, (cudaStream_t)cuda_stream));
  ctx->launches++;
  return PRHF_OK;
}

int prhf_selftest_math(prhf_ctx* ctx, double* max_rel_err6) {
  if (!ctx || !max_rel_err6) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  double* d = nullptr;
  PRHF_CUDA(ctx, cudaMalloc(&d, 6 * sizeof(double)));
  PRHF_CUDA(ctx, cudaMemsetAsync(d, 0, 6 * sizeof(double), ctx->stream));
  cudaError_t e = prhf::launch_math_selftest(1 << 24, d, ctx->stream);
  ctx->launches++;
  double h[6] = {0, 0, 0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, e);
  for (int k = 0; k < 6; ++k) max_rel_err6[k] = h[k];
  return PRHF_OK;
}

int prhf_measure_fp64_peak(prhf_ctx* ctx, double* tflops_out) {
  if (!ctx || !tflops_out) return PRHF_ERR_INVALID_ARG;
  DeviceGuard g(ctx->device);
  double* d = nullptr;
  PRHF_CUDA(ctx, cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  PRHF_CUDA(ctx, cudaEventCreate(&e0));
  PRHF_CUDA(ctx, cudaEventCreate(&e1));
  const int blocks = ctx->sm_count * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    cudaError_t e = prhf::launch_dfma_probe(d, blocks, iters, ctx->stream);
    ctx->launches++;
    cudaEventRecord(e1, ctx->stream);
    if (e != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) {
      cudaFree(d);
      return fail(ctx, e != cudaSuccess ? e : cudaGetLastError());
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return PRHF_OK;
}

}  // extern "C"

/*
 * pyrayhf_b200 -- C ABI of the B200-native vertical forward operator.
 *
 * The reference (victoriyaforsythe/PyRayHF v0.1.0) is pure Python and has no FFI; the seam it
 * offers for this path is the module-level function
 *     PyRayHF.library.vertical_forward_operator(freq, den, bmag, bpsi, alt, mode, n_points)
 * (PyRayHF/library.py:459-509, called by model_VH at library.py:589-591).  The entry points
 * below are what a ctypes binding of that function binds; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns a prhf status code (0 = ok) and
 *     never throws.  prhf_error_string() describes a code.
 *   - units as the reference docstring (library.py:463-478): MHz, m^-3, Tesla, degrees, km.
 *   - mode: 0 = 'O', 1 = 'X' (library.py:391-396).
 *   - a prhf_ctx belongs to one CUDA device and caches the stretched-grid multiplier tables
 *     (one per n_points) plus a small workspace that serves ONE launch sequence at a time.  Calls on
 *     one ctx must come from one host thread at a time; calls that arrive on different streams are
 *     ordered on the device (the later stream waits for the earlier call), they do not overlap.
 *     Create one ctx per host thread / per stream for concurrency (the Python package keeps one per
 *     (device, thread)).
 *   - per-profile status (int32): 0 ok; 1 negative density below the peak (the reference raises
 *     ValueError, library.py:93-94); 2 density peak at index 0 (the reference raises IndexError at
 *     library.py:399).  Rows of a failed profile are NaN.
 */
#ifndef PYRAYHF_B200_H
#define PYRAYHF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRHF_VERSION 100 /* 0.1.0 */

/* status codes */
#define PRHF_OK 0
#define PRHF_ERR_INVALID_ARG 1    /* null pointer, negative size, n_points < 1 ...            */
#define PRHF_ERR_BAD_MODE 2       /* mode not 0/1: ValueError("mode must be 'O' or 'X'")      */
#define PRHF_ERR_CUDA 3           /* a CUDA runtime call failed; see prhf_last_cuda_error()   */
#define PRHF_ERR_NALT_TOO_LARGE 4 /* n_alt exceeds the shared-memory staging limit (regrid stage and tracers \
                                     only; the forward operator switches to a global-memory form) */
#define PRHF_ERR_NO_DEVICE 5      /* no CUDA device / device is not sm_100                    */

/* flags */
#define PRHF_FLAG_DEFAULT 0u
#define PRHF_FLAG_LITERAL 1u /* evaluate library.py:209-254 operation by operation (IEEE div/sqrt, \
                                libdevice sincos) instead of the restructured fast form */

/* per-profile status values */
#define PRHF_PROFILE_OK 0
#define PRHF_PROFILE_NEGATIVE_DENSITY 1
#define PRHF_PROFILE_PEAK_AT_BOTTOM 2

typedef struct prhf_ctx prhf_ctx;

int prhf_version(void);
const char* prhf_error_string(int code);
/* cudaError_t of the last failing runtime call on this ctx (0 if none) and its text. */
int prhf_last_cuda_error(const prhf_ctx* ctx, const char** text);

/* device < 0: the current device. */
int prhf_ctx_create(int device, prhf_ctx** out);
void prhf_ctx_destroy(prhf_ctx* ctx);

/* Largest n_alt the shared-memory staging holds on this device (about 2 400 levels).  prhf_vfo_*_f64 accept longer
 * profiles -- the reference has no limit (np.interp, library.py:424-426) -- through a slower form that keeps the
 * levels in global memory; prhf_regrid_f64 and prhf_snell_f64 return PRHF_ERR_NALT_TOO_LARGE beyond it. */
int prhf_max_n_alt(const prhf_ctx* ctx);

/*
 * Stretched-grid multiplier m[n_points] in [0,1] (replaces smooth_nonuniform_grid(0, 1, n, 10.),
 * library.py:296-321 as called from library.py:361-364).  m_out is a DEVICE pointer.
 * Asynchronous on cuda_stream.
 */
int prhf_grid_multiplier_f64(prhf_ctx* ctx, int n_points, double* m_out, void* cuda_stream);

/*
 * Batched vertical forward operator on DEVICE buffers (replaces library.py:459-509 for every
 * profile of the batch; row p of vh_out equals the reference called on profile p alone).
 *
 *   freq_mhz  [n_freq] when freq_profile_stride == 0, else profile p reads freq_mhz + p*stride
 *   den, bmag, bpsi  [n_profiles x n_alt] row-major
 *   alt       [n_alt] when alt_profile_stride == 0, else profile p reads alt + p*stride
 *   vh_out    [n_profiles x n_freq] row-major, km, NaN where the ray does not reflect
 *   status    [n_profiles] int32 (may be NULL)
 *
 * Asynchronous on cuda_stream; no host synchronisation.  All pointers are device pointers owned
 * by the caller.
 */
int prhf_vfo_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride,
                 const double* den, const double* bmag, const double* bpsi, const double* alt,
                 int64_t alt_profile_stride, int64_t n_profiles, int n_alt, int mode, int n_points,
                 unsigned flags, double* vh_out, int* status, void* cuda_stream);

/*
 * Same operator on HOST buffers: stages the inputs through pinned memory, copies them to the
 * device, runs prhf_vfo_f64 and copies vh/status back; returns after the results are in
 * vh_out / status.  This is the call the Python drop-in makes for numpy inputs.
 */
int prhf_vfo_host_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride,
                      const double* den, const double* bmag, const double* bpsi, const double* alt,
                      int64_t alt_profile_stride, int64_t n_profiles, int n_alt, int mode,
                      int n_points, unsigned flags, double* vh_out, int* status);

/*
 * Same operator, pipelined, for LARGE batches whose buffers live wherever the caller has them (still
 * library.py:459-509 per profile).  Every pointer may be a host pointer (page-locked memory recommended:
 * cudaHostAlloc / cudaHostRegister / torch pin_memory; pageable memory works but copies serialise) or a device
 * pointer, independently of the others; the library looks the pointer up (cudaPointerGetAttributes).  Profiles
 * are processed in chunks of chunk_profiles (0 = automatic) through two device staging slots: the host-to-device
 * copy of chunk k+1, the kernels of chunk k and the device-to-host copy of chunk k-1 run concurrently on three
 * streams, with no staging memcpy on the host.  This is the entry the profile-sharded multi-GPU operator uses:
 * each rank passes its shard and a vh_out that points into one page-locked result buffer shared by all ranks
 * ("final host gather", no collective on the data path).
 *   vh_profile_stride   doubles between consecutive profiles' rows in vh_out (0 = n_freq, dense); a larger
 *                       stride lets interleaved shards land in place in the gathered [P x n_freq] array
 *   cuda_stream         compute stream (NULL = the ctx's own).  Work is ordered after what is already enqueued
 *                       there, and the stream is made to wait for the last copy-out, so "after this call in
 *                       stream order" means "results delivered"
 *   synchronize         non-zero: return after the results are in vh_out / status
 */
int prhf_vfo_stream_f64(prhf_ctx* ctx, const double* freq_mhz, int n_freq, int64_t freq_profile_stride,
                        const double* den, const double* bmag, const double* bpsi, const double* alt,
                        int64_t alt_profile_stride, int64_t n_profiles, int n_alt, int mode, int n_points,
                        unsigned flags, int64_t chunk_profiles, double* vh_out, int64_t vh_profile_stride, int* status,
                        void* cuda_stream, int synchronize);

/* Page-lock (cudaHostRegister, portable) / release a host range owned by the caller, e.g. the POSIX shared-memory
 * segment in which the ranks of one node assemble the [P x n_freq] result of the profile-sharded operator. */
int prhf_host_register(void* ptr, size_t bytes);
int prhf_host_unregister(void* ptr);

/*
 * Elementwise phase / group refractive index on DEVICE buffers (replaces find_mu_mup,
 * library.py:161-256, magnetised branch or isotropic branch chosen by the caller through
 * `isotropic`, because the reference decides it from the whole array, library.py:201).
 * mu_out / mup_out may be NULL.
 */
int prhf_mu_mup_f64(prhf_ctx* ctx, const double* X, const double* Y, const double* bpsi_deg, int64_t n,
                    int mode, int isotropic, unsigned flags, double* mu_out, double* mup_out,
                    void* cuda_stream);

/*
 * The individual stages of the path as standalone operators on DEVICE buffers (the fused operator above
 * never materialises their arrays; these exist because the reference exposes every stage as a public
 * function and its tutorial notebook plots the regridded arrays).  All asynchronous on cuda_stream.
 *
 * prhf_den2freq_f64   replaces den2freq, library.py:75-97: freq_out[i] = sqrt(density[i]) * 8.97866275.
 *                     *negative_flag (device int, may be NULL; caller zeroes it) is OR-ed with 1 if any
 *                     density < 0 -- the reference raises ValueError("Density must be non-negative") there.
 * prhf_find_x_f64     replaces find_X, library.py:120-137: X = den2freq(n_e)**2 / f**2 in that rounding order.
 *                     Strides are 0 (scalar operand) or 1.
 * prhf_find_y_f64     replaces find_Y, library.py:140-158: Y = g_p * b / f.
 * prhf_smooth_grid_f64 replaces smooth_nonuniform_grid(start, end, n_points, sharpness), library.py:296-321.
 */
int prhf_den2freq_f64(prhf_ctx* ctx, const double* density, int64_t n, double* freq_out, int* negative_flag,
                      void* cuda_stream);
int prhf_find_x_f64(prhf_ctx* ctx, const double* n_e, int64_t n_e_stride, const double* f_hz, int64_t f_stride,
                    int64_t n, double* x_out, int* negative_flag, void* cuda_stream);
int prhf_find_y_f64(prhf_ctx* ctx, const double* f_hz, int64_t f_stride, const double* b, int64_t b_stride, int64_t n,
                    double* y_out, void* cuda_stream);
int prhf_smooth_grid_f64(prhf_ctx* ctx, double start, double end, int n_points, double sharpness, double* x_out,
                         void* cuda_stream);

/*
 * Replaces regrid_to_nonuniform_grid, library.py:324-438, for ONE profile: truncation below the density peak,
 * critical curve, reflection heights, stretched altitudes and the profile sampled on them.
 *   f_hz [n_freq] in Hz (this stage of the reference takes Hz, library.py:333-334); n_e, b, bpsi, aalt [n_alt]
 *   crit_height [n_freq]: h_c (library.py:407), NaN where the frequency never reflects (required)
 *   alt_out, dist_out, den_out, bmag_out, bpsi_out [n_freq x n_points] row-major: the 'alt', 'dist', 'den',
 *   'bmag', 'bpsi' entries of the reference's dict (library.py:430-438); any of them may be NULL.
 *   status [1] device int (may be NULL): 0 ok, 1 negative density below the peak (library.py:94),
 *   2 density peak at index 0 (library.py:399).  The 'freq', 'crit_height' and 'ind' entries are broadcasts
 *   the host builds.
 */
int prhf_regrid_f64(prhf_ctx* ctx, const double* f_hz, int n_freq, const double* n_e, const double* b,
                    const double* bpsi, const double* aalt, int n_alt, int mode, int n_points, double* crit_height,
                    double* alt_out, double* dist_out, double* den_out, double* bmag_out, double* bpsi_out, int* status,
                    void* cuda_stream);

/*
 * Replaces find_vh, library.py:259-293: mu' from (X, Y, bpsi) [n_rows x n_cols], vh[r] = nansum(mu' * dh) with
 * 0 -> NaN, + alt_min.  The unmagnetised switch (library.py:201, nanmax|Y| < 1e-12 over the WHOLE array) is
 * evaluated on the device.
 */
int prhf_find_vh_f64(prhf_ctx* ctx, const double* X, const double* Y, const double* bpsi_deg, const double* dh,
                     int64_t n_rows, int64_t n_cols, double alt_min, int mode, unsigned flags, double* vh_out,
                     void* cuda_stream);

/*
 * Synthetic inputs on the DEVICE: two Chapman layers + centred axial dipole (the benchmark generator of
 * pyrayhf_b200/synth.py; stands in for generate_input_1D / calculate_magnetic_field, library.py:2390-2694, whose
 * PyIRI / IGRF dependencies are not available offline).  params [n_profiles x 5] = {foF2 MHz, hmF2 km, scale
 * height km, foE MHz, latitude deg}; alt [n_alt]; den_out, bmag_out, bpsi_out [n_profiles x n_alt].
 */
int prhf_synth_profiles_f64(prhf_ctx* ctx, const double* params, int64_t n_profiles, const double* alt, int n_alt,
                            double* den_out, double* bmag_out, double* bpsi_out, void* cuda_stream);

/*
 * Stratified Snell's-law ray tracers on DEVICE buffers, batched over rays (replace trace_ray_cartesian_snells,
 * library.py:1096-1268 with its helpers library.py:1034-1093, and trace_ray_spherical_snells,
 * library.py:1460-1713; the reference traces one ray per call).  Ray i has frequency f0_hz[i] (Hz) and launch
 * elevation elevation_deg[i]; all rays share the profile alt_km, ne, babs, bpsi [n_alt].
 *   geometry 0 = flat Earth, 1 = spherical Earth (dz_target_km, apex_boost, max_substeps, r_e_km as the
 *   reference's keyword arguments; its defaults are 1.0, 200.0, 400, 6371.0)
 *   scalars_out [n_rays x 5]: group_path_km, group_delay_sec, x_midpoint, z_midpoint, ground_range_km
 *   (the reference's x_apex_km / z_apex_km repeat the midpoint, library.py:1267-1268); NaN when there is no ray
 *   x_out, z_out [n_rays x path_stride] (both or neither; path_stride >= 2 (n_alt + 1) + 1): the ray path,
 *   NaN-padded; n_path_out [n_rays] (may be NULL): points on the path, 0 = no ray.
 */
int prhf_snell_f64(prhf_ctx* ctx, const double* f0_hz, const double* elevation_deg, int64_t n_rays,
                   const double* alt_km, const double* ne, const double* babs, const double* bpsi, int n_alt, int mode,
                   int geometry, unsigned flags, double dz_target_km, double apex_boost, int max_substeps, double r_e_km,
                   double* scalars_out, double* x_out, double* z_out, int path_stride, int* n_path_out,
                   void* cuda_stream);

/*
 * The same tracers for a whole (frequency x elevation) FAN: ray f * n_elev + e has frequency f0_hz[f] and launch
 * elevation elevation_deg[e].  The refractive-index field (find_mu_mup at the profile levels, library.py:1181-1185)
 * depends on the frequency only and is computed once per frequency instead of once per ray; every output is
 * bit-identical to prhf_snell_f64 called with the n_freq * n_elev (frequency, elevation) pairs written out.
 *   scalars_out [n_freq * n_elev x 5], x_out / z_out [n_freq * n_elev x path_stride], n_path_out [n_freq * n_elev].
 */
int prhf_snell_fan_f64(prhf_ctx* ctx, const double* f0_hz, int n_freq, const double* elevation_deg, int n_elev,
                       const double* alt_km, const double* ne, const double* babs, const double* bpsi, int n_alt,
                       int mode, int geometry, unsigned flags, double dz_target_km, double apex_boost, int max_substeps,
                       double r_e_km, double* scalars_out, double* x_out, double* z_out, int path_stride,
                       int* n_path_out, void* cuda_stream);

/*
 * Residual of the inversion objective on DEVICE buffers (replaces the arithmetic tail of residual_VH,
 * library.py:660-668, for a batch of candidate profiles): NaN model heights are replaced by
 * max(nanmean|vh_model[p,:]|, 100) (library.py:664-665), residual = vh_obs - vh_model (library.py:668).
 *   vh_model [n_profiles x n_freq], vh_obs [n_freq]
 *   residual_out [n_profiles x n_freq] (may be NULL), chi2_out [n_profiles] = sum of squared residuals, the
 *   quantity lmfit's brute-force search minimises at library.py:794-798 (may be NULL).
 */
int prhf_residual_f64(prhf_ctx* ctx, const double* vh_model, const double* vh_obs, int64_t n_profiles, int n_freq,
                      double* residual_out, double* chi2_out, void* cuda_stream);

/*
 * Grid node with the smallest objective on DEVICE buffers: the selection step of lmfit's brute-force search as
 * minimize_parameters uses it (library.py:794-798; scipy.optimize.brute takes the argmin of the raveled grid, first
 * minimum wins).  values [n] (e.g. chi2_out of prhf_residual_f64); NaN entries are skipped.
 * out2 [2] (device): {index as a double, -1 when every entry is NaN; the minimum, NaN when none}.
 */
int prhf_argmin_f64(prhf_ctx* ctx, const double* values, int64_t n, double* out2, void* cuda_stream);

/* Device FP64 FMA throughput probe used for the roofline denominator: runs a dependent-free DFMA
 * kernel and returns the measured TFLOP/s (2 flop per FMA). */
int prhf_measure_fp64_peak(prhf_ctx* ctx, double* tflops_out);

/* Accuracy self-test of the kernel's fast reciprocal / reciprocal-square-root primitives: maximum
 * relative error against IEEE division / sqrt over 2^24 samples (log-uniform 1e-30..1e30 and one binade).
 * out[0] rcp_fast, out[1] rsqrt_fast, out[2] raw MUFU.RCP64H seed, out[3] raw MUFU.RSQ64H seed,
 * out[4] / out[5] the seeds after one cubically convergent step. */
int prhf_selftest_math(prhf_ctx* ctx, double* max_rel_err6);

/* Measurement mode for the roofline: while enabled, prhf_vfo_f64 records CUDA events around the row-setup
 * kernel and the tile kernel of every launch pair and synchronises on them (slower: no overlap between the
 * two).  Each call returns the durations accumulated since the previous call, then clears them. */
int prhf_kernel_timing(prhf_ctx* ctx, int enable, double* rows_kernel_ms, double* tile_kernel_ms, int* launch_pairs);

/* Number of kernel launches issued through this ctx since creation (for bench accounting). */
int64_t prhf_launch_count(const prhf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PYRAYHF_B200_H */

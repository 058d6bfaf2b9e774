/*
 * maxpro_b200.h -- C ABI of the B200-native engine for the maxpro hot path.
 *
 * The reference (mrhheffernan/maxpro, pure Rust) has no FFI layer of its own;
 * the boundary this library sits behind is the reference's public surface
 * (SURVEY.md section 8b).  Each entry point below names the reference
 * interface it replaces (paths relative to /root/reference).  INTEGRATION.md
 * shows the Rust `extern "C"` block and the ctypes stub that bind them.
 *
 * Conventions
 *   - Designs are row-major double[n*d]:  design[i*d + k] == Vec<Vec<f64>>[i][k].
 *   - The caller allocates and owns every buffer; nothing is retained after
 *     return.  Plain pointers and sizes only.
 *   - Every function returns a maxpro_status.  On failure a thread-local
 *     message is available from maxpro_last_error(); argument errors reuse the
 *     reference's assert!/PyValueError texts so bindings can re-raise verbatim.
 *   - No C++ exception crosses the boundary.  There is NO CPU fallback: without
 *     a CUDA device every compute entry point returns MAXPRO_ERR_NO_DEVICE.
 *   - Re-entrant and thread-safe; each call uses the calling thread's current
 *     engine device (maxpro_set_device, default 0), or -- build_lhd, search_range and
 *     anneal_lhd -- every device the thread selected with maxpro_set_devices.
 *   - metric: MAXPRO_METRIC_MAXPRO minimises, MAXPRO_METRIC_MAXIMIN maximises
 *     (src/lib.rs:54-63, src/main.rs:65-68).
 */
#ifndef MAXPRO_B200_H
#define MAXPRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAXPRO_API __attribute__((visibility("default")))
#else
#define MAXPRO_API
#endif

typedef enum {
    MAXPRO_OK = 0,
    MAXPRO_ERR_INVALID = 1,     /* argument rejected (reference assert!/ValueError) */
    MAXPRO_ERR_CUDA = 2,        /* CUDA runtime error */
    MAXPRO_ERR_NO_DEVICE = 3,   /* no usable sm_100 device: the engine has no CPU path */
    MAXPRO_ERR_OOM = 4,         /* device or host allocation failed */
    MAXPRO_ERR_UNSUPPORTED = 5  /* shape outside what the kernels cover */
} maxpro_status;

typedef enum {
    MAXPRO_METRIC_NONE = -1,    /* Option::None in build_lhd (src/lib.rs:46-50) */
    MAXPRO_METRIC_MAXPRO = 0,   /* enums::Metrics::MaxPro  (src/enums.rs:5) */
    MAXPRO_METRIC_MAXIMIN = 1   /* enums::Metrics::MaxiMin (src/enums.rs:6) */
} maxpro_metric;

/* flags for maxpro_search_range_device / maxpro_search_range_ex / maxpro_build_lhd_ex.
 * Two-stage search (small designs: any n from 8 to 256 with n*d up to ~1700 and d <= 10; MaxPro from d = 2 up --
 * csrc/search_screen.cu: search_screen_supported): an fp32 screening pass with a proven error bound keeps every
 * candidate that can still be the winner or tie with it, the exact fp64 kernel re-scores those.  Same winner,
 * same score and same tie rule as the fp64-only search (csrc/search_screen.cu derives the bound; the parity tests
 * compare the two paths), at about twice the rate.
 *   MAXPRO_FLAG_DEFAULT      the engine chooses: two-stage for the shapes above when the range holds at least
 *                            32,768 candidates, fp64-only otherwise
 *   MAXPRO_FLAG_FP32_SCREEN  two-stage wherever the shape has a screening kernel, whatever the range size
 *   MAXPRO_FLAG_FP64_ONLY    every candidate scored in fp64 (the kernel the FP64 roofline figures describe) */
#define MAXPRO_FLAG_DEFAULT 0u
#define MAXPRO_FLAG_FP32_SCREEN 1u
#define MAXPRO_FLAG_FP64_ONLY 2u

/* ---- housekeeping ------------------------------------------------------ */
MAXPRO_API const char *maxpro_last_error(void);
MAXPRO_API const char *maxpro_version(void);
/* The design stream generate_lhd / build_lhd draw from (the reference's rand 0.9.2 StdRng stream is not pinned
 * by any reference test, so the engine defines its own): 2 = LHD-mix v2, the default --
 * one Philox4x32-10 call per column for the key, a 32-bit multiply-xorshift mixer per element; 1 = LHD-Philox
 * v1, the stream of the first release (one Philox call per two elements).  Process-wide.  Every kernel, the
 * winner regeneration and the CPU oracle implement both; a seed names different designs under the two.
 * Set it BEFORE asking for workspace sizes: which kernel serves a shape can depend on the stream (under v2 the
 * CTA-per-candidate search keeps designs up to 227 KB on chip, under v1 up to 182 KB), and a device-side call whose
 * workspace was sized under the other version is refused with MAXPRO_ERR_INVALID rather than served. */
MAXPRO_API int maxpro_set_stream_version(int version);
MAXPRO_API int maxpro_get_stream_version(int *version);
MAXPRO_API int maxpro_device_count(int *count);
MAXPRO_API int maxpro_set_device(int device);
MAXPRO_API int maxpro_get_device(int *device);
/* The reference's build_lhd spreads its candidates over every core of the machine through rayon
 * (src/lib.rs:65-98; src/main.rs:71 gets that for free).  The counterpart here: the devices that the
 * host-facing maxpro_build_lhd[_ex], maxpro_search_range and maxpro_anneal_lhd calls of this thread spread
 * their candidates / chains over.  Device g takes the g-th contiguous block of GLOBAL candidate (chain) indices,
 * so the result is the same for any device list; the per-device winners are folded on the host with the
 * reference's tie rule (maxpro_reduce_best).  count == 0 selects every device of the box.  The first device
 * of the list also serves the calls that do not shard (criteria, ordering, winner regeneration).
 * maxpro_set_device(d) is maxpro_set_devices(&d, 1). */
MAXPRO_API int maxpro_set_devices(const int *devices, int count);
/* *count = devices selected; the first min(capacity, *count) go to devices[] (may be NULL). */
MAXPRO_API int maxpro_get_devices(int *devices, int capacity, int *count);
/* Number of engine kernels launched by this process so far (bench.py's gpu_launches). */
MAXPRO_API uint64_t maxpro_kernel_launch_count(void);
/* The library keeps the device blocks of finished calls for the next call of the same host thread: at most
 * 256 MB per thread by default, no block above a quarter of the limit, and only blocks whose stream has nothing
 * pending, so a kept block is never in use.  maxpro_release_device_cache returns them to the driver; *bytes (may
 * be NULL) receives how much was released.  maxpro_set_device_cache_limit changes the per-thread limit for the
 * whole process (0 = keep nothing).  The reference has no counterpart: its Vec<Vec<f64>> temporaries go back to
 * the allocator when a call returns. */
MAXPRO_API int maxpro_release_device_cache(uint64_t *bytes);
MAXPRO_API int maxpro_set_device_cache_limit(uint64_t bytes);
/* Device scratch one maxpro_anneal_lhd call allocates on the current device.  The chains' best designs are folded
 * per CTA on the device, so this grows with the number of RESIDENT chains, not with n_chains (47 MB for 65,536
 * chains of a 1000 x 20 design). */
MAXPRO_API int maxpro_anneal_scratch_bytes(uint64_t n, uint64_t d, int metric, int swap,
                                           uint64_t n_chains, size_t *bytes);

/* ---- reference public functions (host buffers) ------------------------ */

/* lhd::generate_lhd(n_samples, n_dim, &mut StdRng::seed_from_u64(seed))  -- src/lhd.rs:97-132
 * (py: generate_lhd, src/lhd.rs:189-213).  n == 0 writes nothing and returns OK
 * (lhd.rs:98-101); d == 0 is MAXPRO_ERR_INVALID (lhd.rs:106).  Stream = maxpro_set_stream_version,
 * keyed by `seed`; every column is a permutation of the n strata. */
MAXPRO_API int maxpro_generate_lhd(uint64_t n, uint64_t d, uint64_t seed, double *out);

/* Candidates lo..lo+count-1 of the search seeded `seed` (design c = generate_lhd(seed+lo+c),
 * src/lib.rs:70), written as `count` consecutive n*d designs.  Lets a caller (or a test)
 * materialise exactly what the fused search kernels score in place. */
MAXPRO_API int maxpro_generate_batch(uint64_t n, uint64_t d, uint64_t seed, uint64_t lo,
                                     uint64_t count, double *out);

/* maxpro_utils::maxpro_criterion(&design) -- src/maxpro_utils.rs:72-83 */
MAXPRO_API int maxpro_criterion(const double *design, uint64_t n, uint64_t d, double *out);

/* maximin_utils::maximin_criterion(&design) -- src/maximin_utils.rs:54-67 */
MAXPRO_API int maxpro_maximin_criterion(const double *design, uint64_t n, uint64_t d, double *out);

/* maximin_utils::calculate_l2_distance(&a, &b) -- src/maximin_utils.rs:26-45.
 * Host arithmetic (d flops; launching a kernel for it would be absurd). */
MAXPRO_API int maxpro_l2_distance(const double *a, const double *b, uint64_t d, double *out);

/* The map + reduce of build_lhd on caller-supplied candidates -- src/lib.rs:65-98.
 * designs = b consecutive n*d designs; scores[b] (may be NULL); *argbest = index of
 * the winner under the reference tie rule (equal score -> later index, lib.rs:85-97).
 * This is the parity entry point.  * Inputs of 64 MB and more are staged through two pinned buffers of 16 MB (per calling thread, freed by
 * maxpro_release_device_cache) filled by up to eight host threads, so pageable memory reaches the device at 19-27 GB/s
 * instead of the ~10 GB/s of a plain copy; pinned or managed memory is copied directly. */
MAXPRO_API int maxpro_score_batch(const double *designs, uint64_t b, uint64_t n, uint64_t d,
                                  int metric, double *scores, uint64_t *argbest);

/* maxpro::build_lhd(n_samples, n_dim, n_iterations, metric, seed) -- src/lib.rs:30-100
 * (py: build_lhd, src/lib.rs:114-163).  Candidate i is generate_lhd(seed + i)
 * (lib.rs:70); generation and scoring are fused on the device and only
 * (score, index) leave the chip; the winner is regenerated from its index.
 * out_score / out_index may be NULL.  metric == MAXPRO_METRIC_NONE returns
 * generate_lhd(seed) (lib.rs:46-50). */
MAXPRO_API int maxpro_build_lhd(uint64_t n, uint64_t d, uint64_t n_iterations, int metric,
                                uint64_t seed, double *out_design, double *out_score,
                                uint64_t *out_index);
MAXPRO_API int maxpro_build_lhd_ex(uint64_t n, uint64_t d, uint64_t n_iterations, int metric,
                                   uint64_t seed, uint32_t flags, double *out_design,
                                   double *out_score, uint64_t *out_index);

/* One shard of build_lhd: candidates lo..hi-1 of the search seeded `seed`
 * (global indices, so a candidate is the same design at any GPU count).
 * Returns the shard's best (score, global index). */
MAXPRO_API int maxpro_search_range(uint64_t n, uint64_t d, int metric, uint64_t seed,
                                   uint64_t lo, uint64_t hi, double *best_score,
                                   uint64_t *best_index);

MAXPRO_API int maxpro_search_range_ex(uint64_t n, uint64_t d, int metric, uint64_t seed,
                                      uint64_t lo, uint64_t hi, uint32_t flags, double *best_score,
                                      uint64_t *best_index);

/* The fold of src/lib.rs:76-98 over per-shard winners (min-loc / max-loc with the
 * later-index tie rule).  Host arithmetic over `count` 16-byte records. */
MAXPRO_API int maxpro_reduce_best(int metric, const double *scores, const uint64_t *indices,
                                  uint64_t count, double *best_score, uint64_t *best_index);

/* anneal::anneal_lhd(&design, n_iterations, initial_temp, cooling_rate, metric,
 * minimize, seed, swap) -- src/anneal.rs:56-143 (py: src/anneal.rs:161-204).
 * Runs n_chains independent Metropolis chains from `design`; chain c draws from
 * stream seed + c.  n_chains == 1 is the reference's single chain.  Returns the
 * best global-best over chains (never worse than the input, anneal.rs:92-94);
 * out_metric / out_chain may be NULL.  An empty design (n == 0) returns OK
 * untouched (anneal.rs:69-71). */
MAXPRO_API int maxpro_anneal_lhd(const double *design, uint64_t n, uint64_t d,
                                 uint64_t n_iterations, double initial_temp, double cooling_rate,
                                 int metric, int minimize, uint64_t seed, int swap,
                                 uint64_t n_chains, double *out_design, double *out_metric,
                                 uint64_t *out_chain);

/* order::order_design(design, metric, minimize) -- src/order.rs:34-105
 * (py: src/order.rs:161-179).  out_design = reordered rows; perm[m] (may be NULL)
 * = input row placed at position m.  n == 0 or d == 0 returns OK (order.rs:38-46). */
MAXPRO_API int maxpro_order_design(const double *design, uint64_t n, uint64_t d, int metric,
                                   int minimize, double *out_design, uint64_t *perm);

/* ---- device-resident entry points (inputs already in HBM) -------------- */
/* `stream` is a cudaStream_t (NULL = legacy default stream).  Nothing is
 * synchronised; results land in device memory.  Used by bench.py's `value`
 * leg and by callers that keep candidates on the device. */

/* Scratch bytes maxpro_search_range_device needs (device memory). */
MAXPRO_API int maxpro_search_workspace_bytes(uint64_t n, uint64_t d, size_t *bytes);

/* d_result: 16 device bytes {double score; uint64_t index}. */
MAXPRO_API int maxpro_search_range_device(uint64_t n, uint64_t d, int metric, uint64_t seed,
                                          uint64_t lo, uint64_t hi, uint32_t flags,
                                          void *d_result, void *d_workspace,
                                          size_t workspace_bytes, void *stream);

MAXPRO_API int maxpro_score_workspace_bytes(uint64_t b, uint64_t n, uint64_t d, size_t *bytes);

/* d_designs: b*n*d doubles in device memory; d_scores: b doubles; d_result as above. */
MAXPRO_API int maxpro_score_batch_device(const double *d_designs, uint64_t b, uint64_t n,
                                         uint64_t d, int metric, double *d_scores,
                                         void *d_result, void *d_workspace,
                                         size_t workspace_bytes, void *stream);

/* Measures the FP64 FMA-pipe peak of the current device with an independent-DFMA
 * kernel (the roofline denominator; MEASURED_PEAKS.json has no FP64 figure).
 * tflops = 2 * DFMA/s / 1e12 (best of `repeats`). */
MAXPRO_API int maxpro_fp64_peak_probe(int repeats, double *tflops, double *kernel_ms);
/* The same for the FP32 FMA pipe (independent FFMA chains): the denominator for the fp32 screening stage. */
MAXPRO_API int maxpro_fp32_peak_probe(int repeats, double *tflops, double *kernel_ms);

/* ---- test and tooling entry points (not part of the reference's surface) ---- */

/* Forces a kernel path or a launch shape (DESIGN.md section 6a lists the switches); value == NULL clears it.
 * Which kernel serves a call never changes its result -- the parity tests use this to run every kernel on shapes
 * another one normally takes.  MAXPRO_* environment variables are read once, at the first look-up of the process,
 * as initial values; nothing on the call path reads the environment. */
MAXPRO_API int maxpro_debug_set(const char *name, const char *value);

/* Stage 1 of the two-stage search alone, over candidates lo..hi-1: the survivor indices (up to `capacity`; *count
 * is the number stage 1 produced), and, if fp32_sums is given (hi - lo doubles), every candidate's fp32 MaxPro sum
 * in the reference's units.  maxpro_debug_screen_threshold: the survivor threshold (same units) that a candidate
 * with fp32 sum s implies.  The tests check the error bound of search_screen.cu against the oracle with these. */
MAXPRO_API int maxpro_debug_screen(uint64_t n, uint64_t d, uint64_t seed, uint64_t lo, uint64_t hi,
                                   uint64_t *survivors, uint64_t capacity, uint64_t *count, double *fp32_sums);
MAXPRO_API int maxpro_debug_screen_threshold(uint64_t n, uint64_t d, double s, double *threshold);

/* out[i] = the device's pow(x[i], y[i]) as the kernels evaluate (S / pairs)^(1/d): the arithmetic of glibc's pow
 * (maxpro_b200/csrc/pow_glibc.cuh), so that criteria tie exactly where the reference's do. */
MAXPRO_API int maxpro_debug_pow(const double *x, const double *y, uint64_t count, double *out);

#ifdef __cplusplus
}
#endif
#endif /* MAXPRO_B200_H */

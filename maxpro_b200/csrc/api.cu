// C ABI of the engine (include/maxpro_b200.h).  Host-side orchestration only:
// validation with the reference's messages, device buffers, kernel dispatch.
// There is no CPU implementation of any criterion here by design: without a
// usable sm_100 device every compute entry point fails with MAXPRO_ERR_NO_DEVICE.
#include "../../include/maxpro_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "pow_glibc.cuh"

namespace mp {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int k) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }
}  // namespace mp

namespace {

using mp::BestRec;

thread_local std::string g_err;
thread_local int g_device = 0;
// devices a host-facing call of this thread spreads its candidates / chains over (maxpro_set_devices); empty = g_device alone
thread_local std::vector<int> g_devices;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    cudaGetLastError();  // clear the sticky-less error state
    int code = (e == cudaErrorMemoryAllocation) ? MAXPRO_ERR_OOM : MAXPRO_ERR_CUDA;
    return fail(code, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                      \
    do {                                                    \
        cudaError_t e_ = (expr);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr); \
    } while (0)

// Device scratch with a small per-thread cache: the host-facing calls are made in tight
// loops (bench.py's e2e leg, the CLI pipeline) and cudaMalloc/cudaFree take driver-wide
// locks, so the blocks of a finished call are kept for the next call on the same thread and
// device: at most g_cache_limit bytes per thread (256 MB by default, maxpro_set_device_cache_limit),
// no block above a quarter of that.  The largest working set of any call is the annealer's
// 2 x resident-CTAs designs (47 MB at n=1000, d=20).  A block is kept only if the thread's stream
// has nothing pending (a call that returns early on an error may leave work behind: its blocks are
// freed, which synchronises), so a cached block is never in use.
struct CachedBlock { void* p; size_t bytes; int device; };
struct BlockCache {
    std::vector<CachedBlock> free_blocks;
    size_t cached_bytes = 0;
    void drop_all() {
        for (auto& b : free_blocks) cudaFree(b.p);
        free_blocks.clear();
        cached_bytes = 0;
    }
    ~BlockCache() { drop_all(); }
};
thread_local BlockCache g_cache;
std::atomic<unsigned long long> g_cache_limit{256ull << 20};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int device = 0;
    ~DevBuf() {
        if (!p) return;
        const size_t limit = (size_t)g_cache_limit.load(std::memory_order_relaxed);
        if (bytes <= limit / 4 && g_cache.cached_bytes + bytes <= limit && cudaStreamQuery(cudaStreamPerThread) == cudaSuccess) {
            g_cache.free_blocks.push_back({p, bytes, device});
            g_cache.cached_bytes += bytes;
        } else {
            cudaGetLastError();  // cudaErrorNotReady of the query is not an error of the call
            cudaFree(p);
        }
    }
    cudaError_t alloc(size_t want) {
        if (want == 0) want = 1;
        cudaGetDevice(&device);
        auto& fb = g_cache.free_blocks;
        size_t best = fb.size();
        for (size_t i = 0; i < fb.size(); ++i)
            if (fb[i].device == device && fb[i].bytes >= want && fb[i].bytes <= 4 * want + 4096 &&
                (best == fb.size() || fb[i].bytes < fb[best].bytes)) best = i;
        if (best != fb.size()) {
            p = fb[best].p; bytes = fb[best].bytes;
            g_cache.cached_bytes -= bytes;
            fb.erase(fb.begin() + best);
            return cudaSuccess;
        }
        bytes = want;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaErrorMemoryAllocation && !fb.empty()) {
            // the blocks this thread kept may be what is missing: give them back and try once more
            cudaGetLastError();
            g_cache.drop_all();
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) { p = nullptr; bytes = 0; }
        return e;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Large host inputs (maxpro_score_batch).  A cudaMemcpy from pageable memory goes through the driver's own staging at
// 8-11 GB/s, one host thread copying; the kernel behind it needs a hundredth of that time.  So: chunks of kStageChunk bytes
// are copied by several host threads into two pinned staging buffers (kept per calling thread, released with the device
// cache) and each chunk is sent with cudaMemcpyAsync while the next one is being staged.  Memory that is already pinned or
// managed, and anything small, takes the plain copy.  MAXPRO_H2D_STAGED=0 switches the staging off (tests, measurements).
constexpr size_t kStageChunk = 16u << 20;
struct HostStaging {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ready = false;
    cudaError_t init() {
        if (ready) return cudaSuccess;
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaHostAlloc(&buf[i], kStageChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) { release(); return e; }
        }
        ready = true;
        return cudaSuccess;
    }
    void release() {
        for (int i = 0; i < 2; ++i) {
            if (ev[i]) cudaEventDestroy(ev[i]);
            if (buf[i]) cudaFreeHost(buf[i]);
            ev[i] = nullptr; buf[i] = nullptr;
        }
        ready = false;
    }
    ~HostStaging() { release(); }
};
thread_local HostStaging g_staging;

void parallel_memcpy(void* dst, const void* src, size_t bytes, unsigned int nthreads) {
    if (nthreads <= 1 || bytes < (4u << 20)) { std::memcpy(dst, src, bytes); return; }
    const size_t part = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    std::vector<std::thread> workers;
    for (unsigned int t = 1; t < nthreads; ++t) {
        const size_t off = (size_t)t * part;
        if (off >= bytes) break;
        const size_t len = std::min(part, bytes - off);
        workers.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    }
    std::memcpy(dst, src, std::min(part, bytes));
    for (auto& w : workers) w.join();
}

cudaError_t h2d_large(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    bool staged = bytes >= 4 * kStageChunk;
    if (const char* e = mp::cfg("MAXPRO_H2D_STAGED")) staged = atoi(e) != 0 && bytes > kStageChunk;
    if (staged) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) { if (attr.type != cudaMemoryTypeUnregistered) staged = false; }
        else cudaGetLastError();
    }
    if (staged && g_staging.init() != cudaSuccess) { cudaGetLastError(); staged = false; }
    if (!staged) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    unsigned int nthreads = std::thread::hardware_concurrency() / 2;
    nthreads = nthreads < 1 ? 1 : (nthreads > 8 ? 8 : nthreads);
    size_t off = 0;
    for (int c = 0; off < bytes; ++c) {
        const int b = c & 1;
        const size_t len = std::min(kStageChunk, bytes - off);
        // the buffer's last transfer is done (a never-recorded event is complete; a call that failed half-way may have left one pending)
        if (cudaError_t e = cudaEventSynchronize(g_staging.ev[b]); e != cudaSuccess) return e;
        parallel_memcpy(g_staging.buf[b], static_cast<const char*>(src) + off, len, nthreads);
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + off, g_staging.buf[b], len, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(g_staging.ev[b], st);
        if (e != cudaSuccess) return e;
        off += len;
    }
    return cudaSuccess;
}

int sm_count_cached[64];

int ensure_device(int* sms = nullptr) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(MAXPRO_ERR_NO_DEVICE,
                    std::string("no CUDA device available; maxpro_b200 has no CPU path (") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + ")");
    }
    if (g_device < 0 || g_device >= count) return fail(MAXPRO_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(g_device));
    if (g_device < 64 && sm_count_cached[g_device] == 0) {
        int major = 0, n_sm = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, g_device));
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g_device));
        if (major != 10)
            return fail(MAXPRO_ERR_NO_DEVICE, "device is not sm_100 (Blackwell B200); the kernels are built for sm_100a only");
        sm_count_cached[g_device] = n_sm;
    }
    if (sms) *sms = g_device < 64 ? sm_count_cached[g_device] : 148;
    return MAXPRO_OK;
}

bool metric_ok(int m) { return m == MAXPRO_METRIC_MAXPRO || m == MAXPRO_METRIC_MAXIMIN; }

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// ---- small device kernels local to the API -----------------------------------

__global__ void set_result_kernel(BestRec* r, double score, unsigned long long index) {
    r->score = score;
    r->index = index;
}

// total = fold(total, chunk) with chunk-local indices rebased (src/lib.rs:85-97)
__global__ void fold_result_kernel(BestRec* total, const BestRec* chunk, unsigned long long base,
                                   int first, int minimize) {
    double s = chunk->score;
    unsigned long long i = chunk->index == mp::kNoIndex ? mp::kNoIndex : chunk->index + base;
    if (first) { total->score = s; total->index = i; return; }
    bool take = minimize ? mp::better<true>(s, i, total->score, total->index)
                         : mp::better<false>(s, i, total->score, total->index);
    if (take) { total->score = s; total->index = i; }
}

__global__ void init_counter_kernel(unsigned int* c) { c[0] = 0u; c[1] = 0u; }  // done counter, survivor count

// maxpro_debug_pow: the device build of pow_glibc.cuh, element-wise (tests compare it with libm bit for bit)
__global__ void debug_pow_kernel(const double* x, const double* y, unsigned long long count, double* out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = mp::pow_glibc(x[i], y[i]);
}

// ---- search dispatch -----------------------------------------------------------

constexpr size_t kMaxBlockRecs = 4096;
constexpr unsigned long long kScreenMinRange = 1ull << 15;  // MAXPRO_FLAG_DEFAULT: ranges below this go straight to the fp64 kernel
constexpr unsigned int kSurvivorCap = 1u << 20;  // 8 MB; a 2^28-candidate range leaves a few tens of thousands
constexpr size_t kGenericChunkBytes = 48ull << 20;  // below a quarter of the block-cache limit: the chunk buffer is kept between calls

unsigned long long generic_chunk(unsigned long long n, unsigned long long d) {
    unsigned long long per = n * d * sizeof(double);
    unsigned long long b = kGenericChunkBytes / (per ? per : 1);
    if (b < 1) b = 1;
    if (b > (1ull << 16)) b = 1ull << 16;
    return b;
}

// which fused kernel a shape takes (MAXPRO_NO_CTA=1 forces the chunked path; used by the tests)
bool use_warp_kernel(unsigned long long n, unsigned long long d) {
    return !mp::search_small_supported(n, d) && mp::search_warp_supported(n, d) && !mp::cfg("MAXPRO_NO_CTA");
}
bool use_cta_kernel(unsigned long long n, unsigned long long d) {
    return !mp::search_small_supported(n, d) && !use_warp_kernel(n, d) && mp::search_cta_supported(n, d) && !mp::cfg("MAXPRO_NO_CTA");
}

struct SearchWs {
    unsigned int* counter;
    BestRec* block_best;
    BestRec* chunk_result;
    unsigned long long* survivors;  // two-stage search: candidates that passed the fp32 screen (kSurvivorCap entries)
    double* designs;
    double* partial;
    double* scores;
    size_t total;
};

SearchWs carve_search_ws(void* base, unsigned long long n, unsigned long long d) {
    SearchWs w{};
    size_t off = 0;
    char* b = static_cast<char*>(base);
    auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align_up(bytes); return p; };
    w.counter = reinterpret_cast<unsigned int*>(take(256));
    w.block_best = reinterpret_cast<BestRec*>(take(kMaxBlockRecs * sizeof(BestRec)));
    w.chunk_result = reinterpret_cast<BestRec*>(take(256));
    w.survivors = reinterpret_cast<unsigned long long*>(take((size_t)kSurvivorCap * sizeof(unsigned long long)));
    if (!mp::search_small_supported(n, d) && !use_cta_kernel(n, d) && !use_warp_kernel(n, d)) {
        unsigned long long bc = generic_chunk(n, d);
        w.designs = reinterpret_cast<double*>(take(bc * n * d * sizeof(double)));
        w.partial = reinterpret_cast<double*>(take(bc * mp::score_items(n) * sizeof(double)));
        w.scores = reinterpret_cast<double*>(take(bc * sizeof(double)));
    }
    w.total = off;
    return w;
}

int search_range_device_impl(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi,
                             uint32_t flags, void* d_result, void* d_ws, size_t ws_bytes, cudaStream_t stream,
                             int sms) {
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "n_samples must be positive and nonzero");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "n_dim must be positive and nonzero");
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (hi < lo) return fail(MAXPRO_ERR_INVALID, "search range is inverted");
    if (n > (1ull << 18)) return fail(MAXPRO_ERR_UNSUPPORTED, "n_samples above 262144 is outside the stratum-exactness bound of the design streams");
    SearchWs w = carve_search_ws(d_ws, n, d);
    if (ws_bytes < w.total) return fail(MAXPRO_ERR_INVALID, "workspace smaller than maxpro_search_workspace_bytes");
    BestRec* result = static_cast<BestRec*>(d_result);
    const bool minimize = metric == MAXPRO_METRIC_MAXPRO;

    if (hi == lo) {  // identity of the fold (src/lib.rs:77-83)
        set_result_kernel<<<1, 1, 0, stream>>>(result, minimize ? INFINITY : -INFINITY, mp::kNoIndex);
        mp::count_launch();
        CUDA_TRY(cudaGetLastError());
        return MAXPRO_OK;
    }
    if (n < 2) {
        // one point: maxpro == 0.0 (maxpro_utils.rs:75-77), maximin == +INF (maximin_utils.rs:57)
        // for every candidate, so the fold keeps the last index.
        set_result_kernel<<<1, 1, 0, stream>>>(result, minimize ? 0.0 : INFINITY, hi - 1);
        mp::count_launch();
        CUDA_TRY(cudaGetLastError());
        return MAXPRO_OK;
    }
    init_counter_kernel<<<1, 1, 0, stream>>>(w.counter);
    mp::count_launch();
    int cyc_cands = 0, scr_cands = 0, scr_threads = 0;
    const bool want_screen = !(flags & MAXPRO_FLAG_FP64_ONLY) && ((flags & MAXPRO_FLAG_FP32_SCREEN) || hi - lo >= kScreenMinRange);
    const bool exact_cyclic = mp::search_cyclic_supported(n, d, &cyc_cands);
    const bool exact_small = !exact_cyclic && mp::search_small_supported(n, d), exact_warp = !exact_cyclic && !exact_small && use_warp_kernel(n, d);
    if (want_screen && mp::search_screen_supported(n, d, metric, &scr_cands, &scr_threads) && (exact_cyclic || exact_small || exact_warp)) {
        // stage 1: fp32 screening with a proven bound; every candidate that can still win goes to the survivor list
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = seed; L.lo = lo; L.hi = hi; L.flags = flags;
        L.threads = scr_threads; L.group = scr_threads / scr_cands;
        unsigned long long total = hi - lo;
        unsigned long long want = (total + scr_cands - 1) / scr_cands;
        L.grid = (int)(want < (unsigned long long)sms ? want : (unsigned long long)sms);
        L.stream = stream;
        unsigned int cap = kSurvivorCap;
        if (const char* v = mp::cfg("MAXPRO_SCREEN_CAP")) { const long c = atol(v); if (c >= 1 && c < (long)kSurvivorCap) cap = (unsigned int)c; }  // tests: force the overflow path
        CUDA_TRY(mp::launch_search_screen(L, w.survivors, w.counter + 1, cap));
        mp::count_launch();
        // stage 2: the exact fp64 kernel of the shape over the survivors (over lo..hi if the list overflowed), usual fold
        mp::SearchLaunch E{};
        E.n = n; E.d = d; E.metric = metric; E.seed = seed; E.lo = lo; E.hi = hi; E.flags = flags;
        E.block_best = w.block_best; E.done_counter = w.counter; E.result = result;
        E.list = w.survivors; E.list_count = 0; E.list_count_dev = w.counter + 1; E.list_cap = cap;
        E.stream = stream;
        unsigned long long per_cta, max_grid = (unsigned long long)sms;
        if (exact_cyclic) { E.group = 2; E.threads = 2 * cyc_cands; per_cta = (unsigned long long)cyc_cands; }
        else if (exact_small) { mp::search_small_pick_config(n, d, &E.group, &E.threads); per_cta = (unsigned long long)(E.threads / E.group); }
        else { E.group = 0; E.threads = 128; per_cta = 4; max_grid = (unsigned long long)mp::search_warp_grid(n, d, sms); }  // 4 candidates (warps) per CTA
        unsigned long long want2 = (total + per_cta - 1) / per_cta;
        E.grid = (int)(want2 < max_grid ? want2 : max_grid);
        CUDA_TRY(exact_cyclic ? mp::launch_search_cyclic(E) : exact_small ? mp::launch_search_small(E) : mp::launch_search_warp(E));
        mp::count_launch();
        return MAXPRO_OK;
    }
    if (mp::search_cyclic_supported(n, d, &cyc_cands)) {
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = seed; L.lo = lo; L.hi = hi; L.flags = flags;
        L.block_best = w.block_best; L.done_counter = w.counter; L.result = result;
        L.group = 2; L.threads = 2 * cyc_cands;
        unsigned long long total = hi - lo;
        unsigned long long want = (total + cyc_cands - 1) / cyc_cands;
        L.grid = (int)(want < (unsigned long long)sms ? want : (unsigned long long)sms);
        L.stream = stream;
        CUDA_TRY(mp::launch_search_cyclic(L));
        mp::count_launch();
        return MAXPRO_OK;
    }
    if (mp::search_small_supported(n, d)) {
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = seed; L.lo = lo; L.hi = hi; L.flags = flags;
        L.block_best = w.block_best; L.done_counter = w.counter; L.result = result;
        mp::search_small_pick_config(n, d, &L.group, &L.threads);
        unsigned long long total = hi - lo;
        unsigned long long per_cta = (unsigned long long)(L.threads / L.group);
        unsigned long long want = (total + per_cta - 1) / per_cta;
        L.grid = (int)(want < (unsigned long long)sms ? want : (unsigned long long)sms);
        L.stream = stream;
        CUDA_TRY(mp::launch_search_small(L));
        mp::count_launch();
        return MAXPRO_OK;
    }
    if (use_warp_kernel(n, d)) {
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = seed; L.lo = lo; L.hi = hi; L.flags = flags;
        L.block_best = w.block_best; L.done_counter = w.counter; L.result = result;
        unsigned long long total = hi - lo;
        unsigned long long want = (total + 3) / 4;  // 4 candidates (warps) per CTA
        unsigned long long g = (unsigned long long)mp::search_warp_grid(n, d, sms);
        L.grid = (int)(want < g ? want : g);
        L.threads = 128; L.group = 0; L.stream = stream;
        CUDA_TRY(mp::launch_search_warp(L));
        mp::count_launch();
        return MAXPRO_OK;
    }
    if (use_cta_kernel(n, d)) {
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = seed; L.lo = lo; L.hi = hi; L.flags = flags;
        L.block_best = w.block_best; L.done_counter = w.counter; L.result = result;
        unsigned long long total = hi - lo;
        unsigned long long g = (unsigned long long)mp::search_cta_grid(n, d, sms);
        L.grid = (int)(total < g ? total : g);
        L.threads = 128; L.group = 0; L.stream = stream;
        CUDA_TRY(mp::launch_search_cta(L));
        mp::count_launch();
        return MAXPRO_OK;
    }
    // Any other shape: materialise candidates chunk by chunk in HBM and score them with
    // the tiled pair kernel.  Same stream definition, same fold; GPU only.
    const unsigned long long bc = generic_chunk(n, d);
    int first = 1;
    for (unsigned long long c0 = lo; c0 < hi; c0 += bc) {
        unsigned long long cnt = hi - c0 < bc ? hi - c0 : bc;
        CUDA_TRY(mp::launch_generate(n, d, seed, c0, cnt, w.designs, stream));
        mp::ScoreLaunch S{};
        S.designs = w.designs; S.b = cnt; S.n = n; S.d = d; S.metric = metric;
        S.in_cube = 1;  // generate_lhd values lie in [0, 1)
        S.scores = w.scores; S.partial = w.partial; S.block_best = w.block_best;
        S.done_counter = w.counter; S.result = w.chunk_result; S.stream = stream;
        CUDA_TRY(mp::launch_score_batch(S));
        fold_result_kernel<<<1, 1, 0, stream>>>(result, w.chunk_result, c0, first, minimize ? 1 : 0);
        mp::count_launch();
        CUDA_TRY(cudaGetLastError());
        first = 0;
    }
    return MAXPRO_OK;
}

struct ScoreWs {
    unsigned int* counter;
    BestRec* block_best;
    double* partial;
    size_t total;
};

ScoreWs carve_score_ws(void* base, unsigned long long b, unsigned long long n) {
    ScoreWs w{};
    size_t off = 0;
    char* p0 = static_cast<char*>(base);
    auto take = [&](size_t bytes) { char* p = p0 ? p0 + off : nullptr; off += align_up(bytes); return p; };
    w.counter = reinterpret_cast<unsigned int*>(take(256));
    w.block_best = reinterpret_cast<BestRec*>(take(kMaxBlockRecs * sizeof(BestRec)));
    w.partial = reinterpret_cast<double*>(take(b * mp::score_items(n) * sizeof(double)));
    w.total = off;
    return w;
}

int score_batch_device_impl(const double* d_designs, uint64_t b, uint64_t n, uint64_t d, int metric,
                            double* d_scores, void* d_result, void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "Cannot pass an empty design");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "Must have finite, positive number of dimensions");
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (!d_scores) return fail(MAXPRO_ERR_INVALID, "d_scores must not be NULL");
    ScoreWs w = carve_score_ws(d_ws, b, n);
    if (ws_bytes < w.total) return fail(MAXPRO_ERR_INVALID, "workspace smaller than maxpro_score_workspace_bytes");
    if (b == 0) {
        if (d_result) {
            set_result_kernel<<<1, 1, 0, stream>>>(static_cast<BestRec*>(d_result),
                                                  metric == MAXPRO_METRIC_MAXPRO ? INFINITY : -INFINITY, mp::kNoIndex);
            mp::count_launch();
        }
        return MAXPRO_OK;
    }
    init_counter_kernel<<<1, 1, 0, stream>>>(w.counter);
    mp::count_launch();
    int cyc_cands = 0;
    if (mp::search_cyclic_supported(n, d, &cyc_cands) && (reinterpret_cast<uintptr_t>(d_designs) & 15u) == 0 &&
        !mp::cfg("MAXPRO_NO_FAST_SCORE")) {
        // headline shapes: the thread-group kernel of the fused search, fed from HBM instead of Philox
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device);
        BestRec* scratch_result = d_result ? static_cast<BestRec*>(d_result) : w.block_best + (kMaxBlockRecs - 1);
        mp::SearchLaunch L{};
        L.n = n; L.d = d; L.metric = metric; L.seed = 0; L.lo = 0; L.hi = b;
        L.block_best = w.block_best; L.done_counter = w.counter; L.result = scratch_result;
        L.group = 2; L.threads = 2 * cyc_cands;
        unsigned long long want = (b + cyc_cands - 1) / cyc_cands;
        L.grid = (int)(want < (unsigned long long)sms ? want : (unsigned long long)sms);
        L.designs = d_designs; L.scores = d_scores; L.stream = stream;
        CUDA_TRY(mp::launch_search_cyclic(L));
        mp::count_launch();
        return MAXPRO_OK;
    }
    mp::ScoreLaunch S{};
    S.designs = d_designs; S.b = b; S.n = n; S.d = d; S.metric = metric; S.scores = d_scores;
    S.partial = w.partial; S.block_best = w.block_best; S.done_counter = w.counter;
    S.result = static_cast<BestRec*>(d_result); S.stream = stream;
    CUDA_TRY(mp::launch_score_batch(S));
    return MAXPRO_OK;
}

// ---- spreading a host-facing call over the devices of maxpro_set_devices --------------------------
// The reference's build_lhd uses the whole machine through rayon (src/lib.rs:65-98); here the whole machine is
// the selected GPUs.  Candidates and chains are independent units keyed by GLOBAL index, so device g takes the
// g-th contiguous block and the result does not depend on the device count.  One host thread per device (each
// with its own stream and scratch); the per-device winners are 16-byte records folded on the host with the
// reference's tie rule -- no data-path collective, no NCCL inside the library (shard.py keeps the NCCL
// all-gather for the one-process-per-GPU launch under torchrun).
std::vector<int> call_devices() {
    if (g_devices.empty()) return std::vector<int>{g_device};
    return g_devices;
}

template <class F>
int run_on_devices(const std::vector<int>& devs, F&& shard) {  // shard(i, device) -> status; runs with g_device = device
    if (devs.size() == 1) {
        const int saved = g_device;
        g_device = devs[0];
        const int rc = shard(0, devs[0]);
        g_device = saved;
        return rc;
    }
    std::vector<int> rcs(devs.size(), MAXPRO_OK);
    std::vector<std::string> errs(devs.size());
    std::vector<std::thread> workers;
    workers.reserve(devs.size());
    for (size_t i = 0; i < devs.size(); ++i)
        workers.emplace_back([&, i]() {
            g_device = devs[i];
            g_devices.clear();
            rcs[i] = shard((int)i, devs[i]);
            if (rcs[i] != MAXPRO_OK) errs[i] = g_err;
        });
    for (auto& w : workers) w.join();
    for (size_t i = 0; i < devs.size(); ++i)
        if (rcs[i] != MAXPRO_OK) return fail(rcs[i], "device " + std::to_string(devs[i]) + ": " + errs[i]);
    return MAXPRO_OK;
}

// one device: candidates lo..hi-1, winner record to the host
int search_range_one(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi, uint32_t flags,
                     double* best_score, uint64_t* best_index) {
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    size_t ws_bytes = 0;
    if (int rc = maxpro_search_workspace_bytes(n, d, &ws_bytes)) return rc;
    DevBuf dr, dw;
    CUDA_TRY(dr.alloc(sizeof(BestRec)));
    CUDA_TRY(dw.alloc(ws_bytes));
    cudaStream_t st = cudaStreamPerThread;
    if (int rc = search_range_device_impl(n, d, metric, seed, lo, hi, flags, dr.p, dw.p, ws_bytes, st, sms)) return rc;
    BestRec rec;
    CUDA_TRY(cudaMemcpyAsync(&rec, dr.p, sizeof(rec), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *best_score = rec.score;
    *best_index = rec.index;
    return MAXPRO_OK;
}

// all selected devices: index blocks, host fold (src/lib.rs:76-98)
int search_range_sharded(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi, uint32_t flags,
                         double* best_score, uint64_t* best_index) {
    std::vector<int> devs = call_devices();
    const uint64_t total = hi - lo;
    if (devs.size() > total) devs.resize(total ? (size_t)total : 1);
    const uint64_t G = devs.size();
    std::vector<double> scores(G);
    std::vector<uint64_t> indices(G);
    const int rc = run_on_devices(devs, [&](int g, int) {
        // total * g / G without overflow: whole blocks plus the remainder spread over the first shards
        const uint64_t q = total / G, r = total % G;
        const uint64_t a = lo + q * (uint64_t)g + ((uint64_t)g < r ? (uint64_t)g : r);
        const uint64_t b = a + q + ((uint64_t)g < r ? 1 : 0);
        return search_range_one(n, d, metric, seed, a, b, flags, &scores[g], &indices[g]);
    });
    if (rc) return rc;
    return maxpro_reduce_best(metric, scores.data(), indices.data(), G, best_score, best_index);
}

struct AnnealScratch {
    size_t design, raw, best, metrics, improved, cta_chain, cta_buf, jitter;
    size_t total() const { return design + raw + best + metrics + improved + cta_chain + cta_buf + jitter; }
};

// device scratch of one anneal call on the current device (the grid comes from the occupancy of the kernel that serves it)
int anneal_scratch(uint64_t n, uint64_t d, int metric, int swap, uint64_t n_chains, int sms, unsigned int* grid, AnnealScratch* sc) {
    mp::AnnealLaunch L{};
    L.n = n; L.d = d; L.metric = metric; L.swap = swap ? 1 : 0; L.n_chains = n_chains;
    CUDA_TRY(mp::anneal_plan_grid(L, sms, grid));
    const size_t nd = n * d;
    sc->design = nd * sizeof(double);
    sc->raw = (2 * n + 2 + d + 4) * sizeof(double);  // raw0, the nearest-neighbour table (n doubles, n 16-bit indices), d + 3 quantisation parameters
    sc->best = (size_t)*grid * 2 * nd * sizeof(double);
    sc->metrics = n_chains * sizeof(double);
    sc->improved = n_chains * sizeof(unsigned int);
    sc->cta_chain = (size_t)*grid * sizeof(unsigned long long);
    sc->cta_buf = (size_t)*grid * sizeof(int);
    sc->jitter = mp::anneal_jitter_scratch_bytes(n, d, swap ? 1 : 0, n_chains);
    return MAXPRO_OK;
}

// one device: chains 0..n_chains-1 of the stream family `seed` (chain c draws from seed + c)
int anneal_one(const double* design, uint64_t n, uint64_t d, uint64_t n_iterations, double initial_temp, double cooling_rate,
               int metric, int minimize, uint64_t seed, int swap, uint64_t n_chains, double* out_design, double* out_metric,
               uint64_t* out_chain) {
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    const size_t nd = n * d;
    unsigned int grid = 0;
    AnnealScratch sc{};
    if (int rc = anneal_scratch(n, d, metric, swap, n_chains, sms, &grid, &sc)) return rc;
    DevBuf dx, db, dm, di, dr, dcur, dcc, dcb;
    if (sc.jitter) CUDA_TRY(dcur.alloc(sc.jitter));
    CUDA_TRY(dx.alloc(sc.design));
    CUDA_TRY(dr.alloc(sc.raw));
    CUDA_TRY(db.alloc(sc.best));  // per-CTA fold of the chains' bests: 2 designs per resident CTA, not one per chain
    CUDA_TRY(dm.alloc(sc.metrics));
    CUDA_TRY(di.alloc(sc.improved));
    CUDA_TRY(dcc.alloc(sc.cta_chain));
    CUDA_TRY(dcb.alloc(sc.cta_buf));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(cudaMemcpyAsync(dx.p, design, nd * sizeof(double), cudaMemcpyHostToDevice, st));
    mp::AnnealLaunch L{};
    L.design = dx.as<double>(); L.n = n; L.d = d; L.iters = n_iterations; L.seed = seed; L.n_chains = n_chains;
    L.t0 = initial_temp; L.cooling = cooling_rate; L.metric = metric; L.minimize = minimize ? 1 : 0; L.swap = swap ? 1 : 0;
    L.grid = grid; L.best_design = db.as<double>(); L.cta_chain = dcc.as<unsigned long long>(); L.cta_buf = dcb.as<int>();
    L.best_metric = dm.as<double>(); L.improved = di.as<unsigned int>(); L.raw0 = dr.as<double>(); L.stream = st;
    L.rowmin0 = dr.as<double>() + 1; L.rowarg0 = reinterpret_cast<unsigned short*>(dr.as<double>() + 1 + n);
    L.qparams = dr.as<double>() + 2 * n + 2;
    L.cur_scratch = dcur.as<double>();
    CUDA_TRY(mp::launch_anneal(L));
    std::vector<double> metrics(n_chains);
    std::vector<unsigned int> improved(n_chains);
    CUDA_TRY(cudaMemcpyAsync(metrics.data(), dm.p, n_chains * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(improved.data(), di.p, n_chains * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // best chain: strictly better metric wins, the lowest chain index on ties
    uint64_t best = 0;
    for (uint64_t c = 1; c < n_chains; ++c)
        if (minimize ? (metrics[c] < metrics[best]) : (metrics[c] > metrics[best])) best = c;
    if (improved[best]) {
        // chain c ran on CTA c mod grid, which kept it: a chain of that CTA with an equal metric and a lower index
        // would have been picked above instead
        const uint64_t cta = best % grid;
        unsigned long long kept = mp::kNoIndex;
        int buf = 0;
        CUDA_TRY(cudaMemcpyAsync(&kept, dcc.as<unsigned long long>() + cta, sizeof(kept), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&buf, dcb.as<int>() + cta, sizeof(buf), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (kept != best || (buf != 0 && buf != 1)) return fail(MAXPRO_ERR_CUDA, "annealer: the best chain's design was not kept by its CTA");
        CUDA_TRY(cudaMemcpyAsync(out_design, db.as<double>() + (cta * 2 + (uint64_t)buf) * nd, nd * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        std::memcpy(out_design, design, nd * sizeof(double));  // never improved: the input is the global best (anneal.rs:93)
    }
    if (out_metric) *out_metric = metrics[best];
    if (out_chain) *out_chain = best;
    return MAXPRO_OK;
}

int criterion_host(const double* design, uint64_t n, uint64_t d, int metric, double* out) {
    if (!design || !out) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "Cannot pass an empty design");  // maxpro_utils.rs:74, maximin_utils.rs:56
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "Must have finite, positive number of dimensions");  // maxpro_utils.rs:79
    return maxpro_score_batch(design, 1, n, d, metric, out, nullptr);
}

}  // namespace

// =============================== exported ======================================

extern "C" {

const char* maxpro_last_error(void) { return g_err.c_str(); }
const char* maxpro_version(void) { return "maxpro_b200 0.2.0 (sm_100a; design streams: LHD-mix v2 (default), LHD-Philox v1)"; }

int maxpro_set_stream_version(int version) {
    if (version != 1 && version != 2) return fail(MAXPRO_ERR_INVALID, "stream version must be 1 (LHD-Philox v1) or 2 (LHD-mix v2)");
    mp::set_stream_version(version);
    return MAXPRO_OK;
}

int maxpro_get_stream_version(int* version) {
    if (!version) return fail(MAXPRO_ERR_INVALID, "null pointer");
    *version = mp::stream_version();
    return MAXPRO_OK;
}
uint64_t maxpro_kernel_launch_count(void) { return mp::g_launches.load(std::memory_order_relaxed); }

int maxpro_release_device_cache(uint64_t* bytes) {
    const uint64_t released = g_cache.cached_bytes;
    g_cache.drop_all();
    g_staging.release();  // the pinned staging buffers of maxpro_score_batch (host memory: not counted in *bytes)
    if (bytes) *bytes = released;
    return MAXPRO_OK;
}

int maxpro_set_device_cache_limit(uint64_t bytes) {
    g_cache_limit.store(bytes, std::memory_order_relaxed);
    if (g_cache.cached_bytes > bytes) g_cache.drop_all();
    return MAXPRO_OK;
}

int maxpro_debug_set(const char* name, const char* value) {
    if (!name || strncmp(name, "MAXPRO_", 7) != 0) return fail(MAXPRO_ERR_INVALID, "switch names start with MAXPRO_");
    mp::cfg_set(name, value);
    return MAXPRO_OK;
}

int maxpro_device_count(int* count) {
    if (!count) return fail(MAXPRO_ERR_INVALID, "null pointer");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
    *count = c;
    return MAXPRO_OK;
}

int maxpro_set_device(int device) {
    if (device < 0) return fail(MAXPRO_ERR_INVALID, "device index out of range");
    g_device = device;
    g_devices.clear();
    return MAXPRO_OK;
}

int maxpro_set_devices(const int* devices, int count) {
    if (count < 0 || (count > 0 && !devices)) return fail(MAXPRO_ERR_INVALID, "null pointer");
    int available = 0;
    cudaError_t e = cudaGetDeviceCount(&available);
    if (e != cudaSuccess) { cudaGetLastError(); available = 0; }
    std::vector<int> devs;
    if (count == 0) {  // every device of the box
        for (int i = 0; i < available; ++i) devs.push_back(i);
        if (devs.empty()) return fail(MAXPRO_ERR_NO_DEVICE, "no CUDA device available; maxpro_b200 has no CPU path");
    }
    for (int i = 0; i < count; ++i) {
        if (devices[i] < 0 || devices[i] >= available) return fail(MAXPRO_ERR_INVALID, "device index out of range");
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(MAXPRO_ERR_INVALID, "device listed twice");
        devs.push_back(devices[i]);
    }
    g_device = devs[0];
    g_devices = devs;
    return MAXPRO_OK;
}

int maxpro_get_devices(int* devices, int capacity, int* count) {
    if (!count) return fail(MAXPRO_ERR_INVALID, "null pointer");
    const std::vector<int> devs = call_devices();
    *count = (int)devs.size();
    for (int i = 0; devices && i < capacity && i < (int)devs.size(); ++i) devices[i] = devs[i];
    return MAXPRO_OK;
}

int maxpro_get_device(int* device) {
    if (!device) return fail(MAXPRO_ERR_INVALID, "null pointer");
    *device = g_device;
    return MAXPRO_OK;
}

int maxpro_generate_lhd(uint64_t n, uint64_t d, uint64_t seed, double* out) {
    if (n == 0) return MAXPRO_OK;  // lhd.rs:98-101
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "n_dim must be a positive, nonzero integer");  // lhd.rs:106
    if (!out) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n > (1ull << 18)) return fail(MAXPRO_ERR_UNSUPPORTED, "n_samples above 262144 is outside the stratum-exactness bound of the design streams");
    if (int rc = ensure_device()) return rc;
    DevBuf buf;
    CUDA_TRY(buf.alloc(n * d * sizeof(double)));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(mp::launch_generate(n, d, seed, 0, 1, buf.as<double>(), st));
    CUDA_TRY(cudaMemcpyAsync(out, buf.p, n * d * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MAXPRO_OK;
}

// Test / tooling helper: candidates lo..lo+count-1 of the search seeded `seed`.
int maxpro_generate_batch(uint64_t n, uint64_t d, uint64_t seed, uint64_t lo, uint64_t count, double* out) {
    if (n == 0 || count == 0) return MAXPRO_OK;
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "n_dim must be a positive, nonzero integer");
    if (!out) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n > (1ull << 18)) return fail(MAXPRO_ERR_UNSUPPORTED, "n_samples above 262144 is outside the stratum-exactness bound of the design streams");
    if (int rc = ensure_device()) return rc;
    DevBuf buf;
    size_t bytes = count * n * d * sizeof(double);
    CUDA_TRY(buf.alloc(bytes));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(mp::launch_generate(n, d, seed, lo, count, buf.as<double>(), st));
    CUDA_TRY(cudaMemcpyAsync(out, buf.p, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MAXPRO_OK;
}

int maxpro_criterion(const double* design, uint64_t n, uint64_t d, double* out) {
    return criterion_host(design, n, d, MAXPRO_METRIC_MAXPRO, out);
}

int maxpro_maximin_criterion(const double* design, uint64_t n, uint64_t d, double* out) {
    return criterion_host(design, n, d, MAXPRO_METRIC_MAXIMIN, out);
}

int maxpro_l2_distance(const double* a, const double* b, uint64_t d, double* out) {
    if (!a || !b || !out) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "Points must have nonzero length");  // maximin_utils.rs:32
    double acc = 0.0;
    for (uint64_t k = 0; k < d; ++k) {
        volatile double diff = a[k] - b[k];  // volatile: no FMA contraction, Rust never fuses
        volatile double sq = diff * diff;
        acc += sq;
    }
    *out = std::sqrt(acc);
    return MAXPRO_OK;
}

int maxpro_score_batch(const double* designs, uint64_t b, uint64_t n, uint64_t d, int metric,
                       double* scores, uint64_t* argbest) {
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "Cannot pass an empty design");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "Must have finite, positive number of dimensions");
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (b == 0) { if (argbest) *argbest = UINT64_MAX; return MAXPRO_OK; }
    if (!designs) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (int rc = ensure_device()) return rc;
    const size_t bytes = b * n * d * sizeof(double);
    size_t ws_bytes = 0;
    if (int rc = maxpro_score_workspace_bytes(b, n, d, &ws_bytes)) return rc;
    DevBuf dd, ds, dr, dw;
    CUDA_TRY(dd.alloc(bytes));
    CUDA_TRY(ds.alloc(b * sizeof(double)));
    CUDA_TRY(dr.alloc(sizeof(BestRec)));
    CUDA_TRY(dw.alloc(ws_bytes));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(h2d_large(dd.p, designs, bytes, st));
    if (int rc = score_batch_device_impl(dd.as<double>(), b, n, d, metric, ds.as<double>(), dr.p, dw.p, ws_bytes, st)) return rc;
    BestRec rec;
    if (scores) CUDA_TRY(cudaMemcpyAsync(scores, ds.p, b * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&rec, dr.p, sizeof(rec), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (argbest) *argbest = rec.index;
    return MAXPRO_OK;
}

int maxpro_search_range(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi,
                        double* best_score, uint64_t* best_index) {
    return maxpro_search_range_ex(n, d, metric, seed, lo, hi, MAXPRO_FLAG_DEFAULT, best_score, best_index);
}

int maxpro_search_range_ex(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi, uint32_t flags,
                           double* best_score, uint64_t* best_index) {
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "n_samples must be positive and nonzero");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "n_dim must be positive and nonzero");
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (hi < lo) return fail(MAXPRO_ERR_INVALID, "search range is inverted");
    double s = 0.0;
    uint64_t i = 0;
    if (int rc = search_range_sharded(n, d, metric, seed, lo, hi, flags, &s, &i)) return rc;
    if (best_score) *best_score = s;
    if (best_index) *best_index = i;
    return MAXPRO_OK;
}

int maxpro_reduce_best(int metric, const double* scores, const uint64_t* indices, uint64_t count,
                       double* best_score, uint64_t* best_index) {
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (count && (!scores || !indices)) return fail(MAXPRO_ERR_INVALID, "null pointer");
    const bool minimize = metric == MAXPRO_METRIC_MAXPRO;
    double bs = minimize ? INFINITY : -INFINITY;  // lib.rs:77-83
    unsigned long long bi = mp::kNoIndex;
    for (uint64_t r = 0; r < count; ++r) {
        bool take = minimize ? mp::better<true>(scores[r], indices[r], bs, bi)
                             : mp::better<false>(scores[r], indices[r], bs, bi);
        if (take) { bs = scores[r]; bi = indices[r]; }
    }
    if (best_score) *best_score = bs;
    if (best_index) *best_index = bi;
    return MAXPRO_OK;
}

int maxpro_build_lhd_ex(uint64_t n, uint64_t d, uint64_t n_iterations, int metric, uint64_t seed,
                        uint32_t flags, double* out_design, double* out_score, uint64_t* out_index) {
    if (n == 0) return fail(MAXPRO_ERR_INVALID, "n_samples must be positive and nonzero");        // lib.rs:37
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "n_dim must be positive and nonzero");            // lib.rs:38
    if (n_iterations == 0) return fail(MAXPRO_ERR_INVALID, "n_iterations must be positive and nonzero");  // lib.rs:39-42
    if (metric != MAXPRO_METRIC_NONE && !metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Metric must be 'maxpro' or 'maximin'.");
    if (!out_design) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (metric == MAXPRO_METRIC_NONE) {  // lib.rs:46-50
        if (out_score) *out_score = NAN;
        if (out_index) *out_index = 0;
        return maxpro_generate_lhd(n, d, seed, out_design);
    }
    BestRec rec;
    if (int rc = search_range_sharded(n, d, metric, seed, 0, n_iterations, flags, &rec.score, reinterpret_cast<uint64_t*>(&rec.index))) return rc;
    if (rec.index == mp::kNoIndex) return fail(MAXPRO_ERR_CUDA, "search produced no winner (all scores NaN?)");
    // only (score, index) left the chips: rebuild the winner from its index (lib.rs:70)
    if (int rc = ensure_device()) return rc;
    DevBuf dx;
    CUDA_TRY(dx.alloc(n * d * sizeof(double)));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(mp::launch_generate(n, d, seed, rec.index, 1, dx.as<double>(), st));
    CUDA_TRY(cudaMemcpyAsync(out_design, dx.p, n * d * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (out_score) *out_score = rec.score;
    if (out_index) *out_index = rec.index;
    return MAXPRO_OK;
}

int maxpro_build_lhd(uint64_t n, uint64_t d, uint64_t n_iterations, int metric, uint64_t seed,
                     double* out_design, double* out_score, uint64_t* out_index) {
    return maxpro_build_lhd_ex(n, d, n_iterations, metric, seed, MAXPRO_FLAG_DEFAULT, out_design, out_score, out_index);
}

int maxpro_anneal_lhd(const double* design, uint64_t n, uint64_t d, uint64_t n_iterations,
                      double initial_temp, double cooling_rate, int metric, int minimize, uint64_t seed,
                      int swap, uint64_t n_chains, double* out_design, double* out_metric, uint64_t* out_chain) {
    if (n == 0) return MAXPRO_OK;  // anneal.rs:69-71
    if (n_iterations == 0) return fail(MAXPRO_ERR_INVALID, "n_iterations must be positive and nonzero");  // anneal.rs:72-75
    if (!(initial_temp > 0.0)) return fail(MAXPRO_ERR_INVALID, "initial_temp must be positive and nonzero");  // anneal.rs:76-79
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Unknown metric. Available metrics are 'maxpro' and 'maximin'.");
    if (d == 0) return fail(MAXPRO_ERR_INVALID, "Must have finite, positive number of dimensions");
    if (n_chains == 0) return fail(MAXPRO_ERR_INVALID, "n_chains must be positive and nonzero");
    if (!design || !out_design) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (!mp::anneal_supported(n, d, swap))
        return fail(MAXPRO_ERR_UNSUPPORTED, "design too large for the annealer (at most 65535 points and ~14000 dimensions)");
    std::vector<int> devs = call_devices();
    if (devs.size() > n_chains) devs.resize((size_t)n_chains);
    if (devs.size() == 1)
        return run_on_devices(devs, [&](int, int) {
            return anneal_one(design, n, d, n_iterations, initial_temp, cooling_rate, metric, minimize, seed, swap, n_chains,
                              out_design, out_metric, out_chain);
        });
    // chain blocks over the devices: chain c draws from stream seed + c wherever it runs
    const uint64_t G = devs.size(), nd = n * d;
    std::vector<std::vector<double>> designs(G, std::vector<double>(nd));
    std::vector<double> metrics(G);
    std::vector<uint64_t> chains(G), first(G);
    const int rc = run_on_devices(devs, [&](int g, int) {
        const uint64_t q = n_chains / G, r = n_chains % G;
        const uint64_t c0 = q * (uint64_t)g + ((uint64_t)g < r ? (uint64_t)g : r), cnt = q + ((uint64_t)g < r ? 1 : 0);
        first[g] = c0;
        return anneal_one(design, n, d, n_iterations, initial_temp, cooling_rate, metric, minimize, seed + c0, swap, cnt,
                          designs[g].data(), &metrics[g], &chains[g]);
    });
    if (rc) return rc;
    uint64_t best = 0;  // strictly better wins: the lowest global chain index on ties (blocks are in chain order)
    for (uint64_t g = 1; g < G; ++g)
        if (minimize ? (metrics[g] < metrics[best]) : (metrics[g] > metrics[best])) best = g;
    std::memcpy(out_design, designs[best].data(), nd * sizeof(double));
    if (out_metric) *out_metric = metrics[best];
    if (out_chain) *out_chain = first[best] + chains[best];
    return MAXPRO_OK;
}

int maxpro_order_design(const double* design, uint64_t n, uint64_t d, int metric, int minimize,
                        double* out_design, uint64_t* perm) {
    if (n == 0 || d == 0) return MAXPRO_OK;  // order.rs:38-46
    if (!metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "Unknown metric. Available metrics are 'maxpro' and 'maximin'.");
    if (!design || !out_design) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n >= (1ull << 32) - 1) return fail(MAXPRO_ERR_UNSUPPORTED, "n_samples too large to index");
    if (int rc = ensure_device()) return rc;
    DevBuf dx, dp, da, dpool;
    CUDA_TRY(dx.alloc(n * d * sizeof(double)));
    CUDA_TRY(dp.alloc(n * sizeof(unsigned long long)));
    CUDA_TRY(da.alloc(n * sizeof(double)));
    CUDA_TRY(dpool.alloc(n * sizeof(unsigned int)));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(cudaMemcpyAsync(dx.p, design, n * d * sizeof(double), cudaMemcpyHostToDevice, st));
    mp::OrderLaunch L{};
    L.design = dx.as<double>(); L.n = n; L.d = d; L.metric = metric; L.minimize = minimize ? 1 : 0;
    L.perm = dp.as<unsigned long long>(); L.acc = da.as<double>(); L.pool = dpool.as<unsigned int>(); L.stream = st;
    const bool no_cluster = mp::cfg("MAXPRO_NO_ORDER_CLUSTER") != nullptr;
    if (!no_cluster && mp::order_cluster_async_supported(n, d) && !mp::cfg("MAXPRO_ORDER_CLUSTER_SYNC")) CUDA_TRY(mp::launch_order_cluster_async(L));
    else if (!no_cluster && mp::order_cluster_supported(n, d)) CUDA_TRY(mp::launch_order_cluster(L));
    else CUDA_TRY(mp::launch_order(L));
    std::vector<unsigned long long> hp(n);
    CUDA_TRY(cudaMemcpyAsync(hp.data(), dp.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<double> tmp;
    const double* src = design;
    if (out_design == design) { tmp.assign(design, design + n * d); src = tmp.data(); }
    for (uint64_t m = 0; m < n; ++m) {
        std::memcpy(out_design + m * d, src + hp[m] * d, d * sizeof(double));
        if (perm) perm[m] = hp[m];
    }
    return MAXPRO_OK;
}

int maxpro_search_workspace_bytes(uint64_t n, uint64_t d, size_t* bytes) {
    if (!bytes) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n == 0 || d == 0) return fail(MAXPRO_ERR_INVALID, "n_samples and n_dim must be positive and nonzero");
    *bytes = carve_search_ws(nullptr, n, d).total;
    return MAXPRO_OK;
}

int maxpro_search_range_device(uint64_t n, uint64_t d, int metric, uint64_t seed, uint64_t lo, uint64_t hi,
                               uint32_t flags, void* d_result, void* d_workspace, size_t workspace_bytes,
                               void* stream) {
    if (!d_result || !d_workspace) return fail(MAXPRO_ERR_INVALID, "null pointer");
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    return search_range_device_impl(n, d, metric, seed, lo, hi, flags, d_result, d_workspace, workspace_bytes,
                                    static_cast<cudaStream_t>(stream), sms);
}

int maxpro_score_workspace_bytes(uint64_t b, uint64_t n, uint64_t d, size_t* bytes) {
    if (!bytes) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n == 0 || d == 0) return fail(MAXPRO_ERR_INVALID, "Cannot pass an empty design");
    *bytes = carve_score_ws(nullptr, b, n).total;
    return MAXPRO_OK;
}

int maxpro_score_batch_device(const double* d_designs, uint64_t b, uint64_t n, uint64_t d, int metric,
                              double* d_scores, void* d_result, void* d_workspace, size_t workspace_bytes,
                              void* stream) {
    if (b && !d_designs) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (!d_workspace) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (int rc = ensure_device()) return rc;
    return score_batch_device_impl(d_designs, b, n, d, metric, d_scores, d_result, d_workspace, workspace_bytes,
                                   static_cast<cudaStream_t>(stream));
}

int maxpro_fp64_peak_probe(int repeats, double* tflops, double* kernel_ms) {
    if (!tflops) return fail(MAXPRO_ERR_INVALID, "null pointer");
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    if (repeats < 1) repeats = 1;
    DevBuf sink;
    CUDA_TRY(sink.alloc(17 * sizeof(double)));
    cudaStream_t st = cudaStreamPerThread;
    double vals[17];
    vals[0] = 0.0;
    for (int i = 0; i < 16; ++i) vals[i + 1] = 0.999999 + 1e-7 * i;
    CUDA_TRY(cudaMemcpyAsync(sink.p, vals, sizeof(vals), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int threads = 256, grid = sms * 4, iters = 1 << 16;
    double best_ms = 1e30;
    for (int r = 0; r < repeats + 1; ++r) {  // first launch is warm-up
        cudaEventRecord(e0, st);
        cudaError_t e = mp::launch_dfma_probe(sink.as<double>(), grid, threads, iters, st);
        cudaEventRecord(e1, st);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return cuda_fail(e, "dfma probe"); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)grid;
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    if (kernel_ms) *kernel_ms = best_ms;
    return MAXPRO_OK;
}

int maxpro_fp32_peak_probe(int repeats, double* tflops, double* kernel_ms) {
    if (!tflops) return fail(MAXPRO_ERR_INVALID, "null pointer");
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    if (repeats < 1) repeats = 1;
    DevBuf sink;
    CUDA_TRY(sink.alloc(17 * sizeof(float)));
    cudaStream_t st = cudaStreamPerThread;
    float vals[17];
    vals[0] = 0.f;
    for (int i = 0; i < 16; ++i) vals[i + 1] = 0.9999f + 1e-5f * (float)i;
    CUDA_TRY(cudaMemcpyAsync(sink.p, vals, sizeof(vals), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int threads = 256, grid = sms * 8, iters = 1 << 16;
    double best_ms = 1e30;
    for (int r = 0; r < repeats + 1; ++r) {  // first launch is warm-up
        cudaEventRecord(e0, st);
        cudaError_t e = mp::launch_ffma_probe(sink.as<float>(), grid, threads, iters, st);
        cudaEventRecord(e1, st);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return cuda_fail(e, "ffma probe"); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)grid;
    *tflops = flops / (best_ms * 1e-3) / 1e12;
    if (kernel_ms) *kernel_ms = best_ms;
    return MAXPRO_OK;
}

int maxpro_anneal_scratch_bytes(uint64_t n, uint64_t d, int metric, int swap, uint64_t n_chains, size_t* bytes) {
    if (!bytes) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (n == 0 || d == 0 || n_chains == 0 || !metric_ok(metric)) return fail(MAXPRO_ERR_INVALID, "n_samples, n_dim and n_chains must be positive and nonzero");
    int sms = 0;
    if (int rc = ensure_device(&sms)) return rc;
    if (!mp::anneal_supported(n, d, swap)) return fail(MAXPRO_ERR_UNSUPPORTED, "design too large for the annealer (at most 65535 points and ~14000 dimensions)");
    unsigned int grid = 0;
    AnnealScratch sc{};
    if (int rc = anneal_scratch(n, d, metric, swap, n_chains, sms, &grid, &sc)) return rc;
    *bytes = sc.total();
    return MAXPRO_OK;
}

int maxpro_debug_screen(uint64_t n, uint64_t d, uint64_t seed, uint64_t lo, uint64_t hi, uint64_t* survivors,
                        uint64_t capacity, uint64_t* count, double* fp32_sums) {
    if (!count) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (hi < lo) return fail(MAXPRO_ERR_INVALID, "search range is inverted");
    int sms = 0, cands = 0, threads = 0;
    if (int rc = ensure_device(&sms)) return rc;
    if (!mp::search_screen_supported(n, d, MAXPRO_METRIC_MAXPRO, &cands, &threads))
        return fail(MAXPRO_ERR_UNSUPPORTED, "no MaxPro screening kernel for this shape (csrc/search_screen.cu: search_screen_supported)");
    const uint64_t total = hi - lo;
    DevBuf dl, dc, ds;
    CUDA_TRY(dl.alloc((size_t)kSurvivorCap * sizeof(unsigned long long)));
    CUDA_TRY(dc.alloc(2 * sizeof(unsigned int)));
    if (fp32_sums) CUDA_TRY(ds.alloc(total * sizeof(double)));
    cudaStream_t st = cudaStreamPerThread;
    init_counter_kernel<<<1, 1, 0, st>>>(dc.as<unsigned int>());
    mp::count_launch();
    mp::SearchLaunch L{};
    L.n = n; L.d = d; L.metric = MAXPRO_METRIC_MAXPRO; L.seed = seed; L.lo = lo; L.hi = hi;
    L.threads = threads; L.group = threads / cands; L.stream = st;
    const uint64_t want = (total + cands - 1) / cands;
    L.grid = (int)(want < (uint64_t)sms ? (want ? want : 1) : (uint64_t)sms);
    CUDA_TRY(mp::launch_search_screen(L, dl.as<unsigned long long>(), dc.as<unsigned int>() + 1, kSurvivorCap,
                                      fp32_sums ? ds.as<double>() : nullptr));
    mp::count_launch();
    unsigned int got = 0;
    CUDA_TRY(cudaMemcpyAsync(&got, dc.as<unsigned int>() + 1, sizeof(got), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *count = got;
    const uint64_t copy = std::min<uint64_t>(std::min<uint64_t>(got, kSurvivorCap), capacity);
    if (survivors && copy) CUDA_TRY(cudaMemcpyAsync(survivors, dl.p, copy * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    if (fp32_sums && total) {
        CUDA_TRY(cudaMemcpyAsync(fp32_sums, ds.p, total * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        double g2 = 1.0;
        mp::screen_debug_constants(n, d, &g2, nullptr, 1.0);
        for (uint64_t t = 0; t < total; ++t) fp32_sums[t] *= g2;  // scaled units -> the reference's S
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return MAXPRO_OK;
}

int maxpro_debug_screen_threshold(uint64_t n, uint64_t d, double s, double* threshold) {
    if (!threshold) return fail(MAXPRO_ERR_INVALID, "null pointer");
    double ratio = 0.0;
    mp::screen_debug_constants(n, d, nullptr, &ratio, s);
    *threshold = s * ratio;
    return MAXPRO_OK;
}

int maxpro_debug_pow(const double* x, const double* y, uint64_t count, double* out) {
    if (count == 0) return MAXPRO_OK;
    if (!x || !y || !out) return fail(MAXPRO_ERR_INVALID, "null pointer");
    if (int rc = ensure_device()) return rc;
    DevBuf dx, dy, dz;
    CUDA_TRY(dx.alloc(count * sizeof(double)));
    CUDA_TRY(dy.alloc(count * sizeof(double)));
    CUDA_TRY(dz.alloc(count * sizeof(double)));
    cudaStream_t st = cudaStreamPerThread;
    CUDA_TRY(cudaMemcpyAsync(dx.p, x, count * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dy.p, y, count * sizeof(double), cudaMemcpyHostToDevice, st));
    debug_pow_kernel<<<(unsigned int)((count + 255) / 256 < 65535 ? (count + 255) / 256 : 65535), 256, 0, st>>>(dx.as<double>(), dy.as<double>(), count, dz.as<double>());
    mp::count_launch();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dz.p, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return MAXPRO_OK;
}

}  // extern "C"

// K1+K2 fused, "large design" shape: one CTA per candidate (e.g. n=500, d=10: 40 KB).
//
// Replaces the same reference code as the other search kernels: generate_lhd
// (src/lhd.rs:97-132) -> criterion (src/maxpro_utils.rs:28-83 / src/maximin_utils.rs:54-67)
// -> fold (src/lib.rs:76-98), for designs too big for the thread-group kernels.  The
// candidate lives in shared memory, dimension-major X[k][i] (row stride: cta_row_stride); nothing but
// one 16-byte record per CTA reaches HBM.
//
// Generation (LHD-Philox v1, bit-identical to generate.cu / the oracle):
//   phase 1  every thread runs Philox for its share of the (column, element-pair) blocks and
//            stages v_i at a[i] and the Fisher-Yates index j_i in a 16-bit side array
//            (the inside-out shuffle never touches a[i] before step i, so a[i] is free);
//   phase 2  one lane per column walks the n dependent steps  t=a[j]; a[i]=t; a[j]=v_i.
// Scoring: 64-row tiles (lane <-> rows i, i+32) against blocks of 8 columns whose coordinates
// are warp-wide broadcasts; the 16 running products per lane stay in registers across the d
// dimensions (2 per-lane + 4 broadcast LDS.128-able loads per dimension for 16 pairs).
// Work items (row tile, column block) are dealt round-robin to the warps; blocks that touch
// the diagonal or the edge take a masked path.  MaxPro terms go through the reciprocal chains
// of common.cuh (candidates are in the unit cube by construction).
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"
#include "tile_score.cuh"

#include <cstdlib>
#include <type_traits>

namespace mp {

struct CtaParams {
    int gen;       // design stream version (philox.cuh)
    int n, d, ns;  // ns = row stride of the [k][i] layout (cta_row_stride)
    unsigned long long seed, lo, hi;
    double c1, c0, n_pairs, inv_d;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

// Row stride of the [k][i] layout: n rounded up to a multiple of 8 (whole partner blocks), plus 2.
// Even, so partner blocks are 16-byte aligned; 2 or 10 modulo 16, so the lanes that walk one
// column each in generation phase 2 (stride ns doubles) fall into 8 different bank groups.
__host__ __device__ __forceinline__ unsigned long long cta_row_stride(unsigned long long n) { return (n + 7) / 8 * 8 + 2; }

// Generation phase 1 for one candidate: every thread runs Philox for its share of the (column,
// element-pair) blocks and stages v_i at a[i] and the Fisher-Yates index j_i in the side array.
__device__ __forceinline__ void cta_stage_candidate(int gen, double* xs, unsigned short* jst, int n, int d, int ns, uint32_t s_lo, uint32_t s_hi,
                                                    double c1, double c0, int tid, int nthreads) {
    const int nb = (n + 1) / 2;
    LhdKeyCache kc{0xFFFFFFFFu, 0u, 0u};
    // block q <-> (column k, element pair b), walked without a division
    int k = 0, b = tid;
    while (b >= nb && k < d) { b -= nb; ++k; }
    while (k < d) {
        const LhdBlock w = lhd_draw_block(gen, (uint32_t)b, (uint32_t)k, s_lo, s_hi, kc);
        const int i = 2 * b;
        xs[k * ns + i] = lhd_value(gen, (uint32_t)i, w.jit0, c1, c0);
        jst[k * ns + i] = (unsigned short)w.j0;
        if (i + 1 < n) {
            xs[k * ns + i + 1] = lhd_value(gen, (uint32_t)i + 1u, w.jit1, c1, c0);
            jst[k * ns + i + 1] = (unsigned short)w.j1;
        }
        b += nthreads;
        while (b >= nb && k < d) { b -= nb; ++k; }
    }
}

// Generation phase 2: lane k of the calling group walks the n dependent steps of column k.
__device__ __forceinline__ void cta_walk_columns(double* xs, const unsigned short* jst, int n, int d, int ns, int first, int step) {
    for (int k = first; k < d; k += step) fy_walk_column(xs + k * ns, jst + k * ns, n);
}

// THREADS: 128 while four CTAs fit one SM; a design that leaves room for two (one) CTAs gets 256 (512) threads -- the same
// 16 warps per SM: one CTA of four warps is one warp per sub-partition, and that does not fill the FP64 pipe.
// COMPACT (stream v2 only): no side array -- the lane that walks a column draws index and value of every step itself (a
// handful of integer operations and one DFMA per step, hidden behind the walk's dependent shared-memory round trip), so a
// design takes 8 bytes per element instead of 10 and designs up to 227 KB stay on this kernel instead of going through HBM.
template <bool MAXPRO, int THREADS, bool COMPACT>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) search_cta_kernel(CtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, d = p.d, ns = p.ns;
    double* xs = reinterpret_cast<double*>(smem_raw);                         // [d][ns]
    unsigned short* jst = reinterpret_cast<unsigned short*>(xs + (size_t)d * ns);  // [d][ns] (not COMPACT)
    double* red = reinterpret_cast<double*>(smem_raw + (((size_t)d * ns * (COMPACT ? 8 : 10) + 15) / 16) * 16);  // [THREADS / 32] warp partials
    __shared__ BestRec warp_best[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = (n + kTileRows - 1) / kTileRows;
    const int chunk = cta_dim_chunk(d);

    LaneBest<MAXPRO> lb;  // fold of src/lib.rs:85-97 on the criterion (reduce.cuh); kept by thread 0
    lb.init(p.inv_d);

    for (unsigned long long idx = p.lo + blockIdx.x; idx < p.hi; idx += gridDim.x) {
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        __syncthreads();  // previous candidate fully consumed
        if (COMPACT) {
            for (int k = tid; k < d; k += THREADS) {
                uint32_t ka, kb;
                lhd2_column_key((uint32_t)k, s_lo, s_hi, ka, kb);
                fy_walk_column_v2(xs + k * ns, ka, kb, n, p.c1, p.c0);
            }
        } else {
            cta_stage_candidate(p.gen, xs, jst, n, d, ns, s_lo, s_hi, p.c1, p.c0, tid, THREADS);
            __syncthreads();
            cta_walk_columns(xs, jst, n, d, ns, tid, THREADS);
        }
        __syncthreads();

        // ---- scoring: items (row tile, column block) dealt round-robin to the warps
        CtaScore st;
        st.reset();
        int t = 0, jb = warp;
        int nblk = (n + kJB - 1) / kJB;  // column blocks from the tile's first row
        while (true) {
            while (t < tiles && jb >= nblk) { jb -= nblk; ++t; nblk = (n - t * kTileRows + kJB - 1) / kJB; }
            if (t >= tiles) break;
            cta_score_item_any<MAXPRO>(chunk, xs, n, d, ns, t * kTileRows, t * kTileRows + jb * kJB, lane, st);
            jb += THREADS / 32;
        }
        if (MAXPRO) st.flush();
        const double acc = st.acc, mn = st.mn;
        double part = MAXPRO ? warp_reduce_sum(acc) : warp_reduce_min(mn);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            double s = red[0];
#pragma unroll
            for (int w = 1; w < THREADS / 32; ++w) s = MAXPRO ? s + red[w] : fmin(s, red[w]);
            if (!MAXPRO) s = sqrt(s);
            lb.offer(s, idx, p.n_pairs, p.inv_d);
        }
    }
    block_publish_best<MAXPRO>(lb.best, lb.idx, warp_best, p.block_best, p.done_counter, p.result);
}

// ---- pipelined variant ---------------------------------------------------------------------------
// Two design buffers per CTA.  While the CTA scores candidate c out of one buffer, candidate c+1 is
// built in the other: every thread runs its share of the Philox blocks at the top of the
// iteration (a few hundred instructions), then ONE warp walks the sequential Fisher-Yates steps
// (n dependent shared-memory round trips, ~20k cycles at n = 500) while the other warps are
// already scoring; it joins them when the walk is done.  Items are handed out by a shared counter,
// so the late warp simply takes fewer.  In the single-buffer kernel above the walk holds three of
// the four warps at a barrier (16% of all warp samples at n = 500, d = 10).
constexpr int kPipeThreads = 256;

// COMPACT (stream v2): nothing is staged -- the walking warp draws index and value of every step itself (fy_walk_column_v2).
template <bool MAXPRO, bool COMPACT>
__global__ void __launch_bounds__(kPipeThreads, 2) search_cta_pipe_kernel(CtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, d = p.d, ns = p.ns;
    const size_t xbytes = (size_t)d * ns * sizeof(double);
    double* const xbase = reinterpret_cast<double*>(smem_raw);
    const size_t xstride = (size_t)d * ns;  // doubles per buffer
    unsigned short* jst = reinterpret_cast<unsigned short*>(smem_raw + 2 * xbytes);  // [d][ns]
    constexpr int kWarps = kPipeThreads / 32;
    __shared__ double red[kWarps];
    __shared__ int next_item;
    __shared__ BestRec warp_best[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = (n + kTileRows - 1) / kTileRows;
    const int chunk = cta_dim_chunk(d);

    LaneBest<MAXPRO> lb;  // fold of src/lib.rs:85-97 on the criterion (reduce.cuh); kept by thread 0
    lb.init(p.inv_d);

    unsigned long long idx = p.lo + blockIdx.x;
    if (idx < p.hi) {  // prologue: the first candidate, nothing to overlap it with
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        if (COMPACT) {
            for (int k = tid; k < d; k += kPipeThreads) {
                uint32_t ka, kb;
                lhd2_column_key((uint32_t)k, (uint32_t)stream, (uint32_t)(stream >> 32), ka, kb);
                fy_walk_column_v2(xbase + k * ns, ka, kb, n, p.c1, p.c0);
            }
        } else {
            cta_stage_candidate(p.gen, xbase, jst, n, d, ns, (uint32_t)stream, (uint32_t)(stream >> 32), p.c1, p.c0, tid, kPipeThreads);
            __syncthreads();
            cta_walk_columns(xbase, jst, n, d, ns, tid, kPipeThreads);
        }
    }
    __syncthreads();  // the side array is rewritten at the top of the loop
    int cur = 0;
    for (; idx < p.hi; idx += gridDim.x, cur ^= 1) {
        const unsigned long long nidx = idx + gridDim.x;
        const bool has_next = nidx < p.hi;
        // the side array is free (the previous walk ended before the last barrier) and so is the other buffer
        if (has_next && !COMPACT) {
            const unsigned long long stream = p.seed + nidx;
            cta_stage_candidate(p.gen, xbase + (cur ^ 1) * xstride, jst, n, d, ns, (uint32_t)stream, (uint32_t)(stream >> 32), p.c1, p.c0, tid, kPipeThreads);
        }
        if (tid == 0) next_item = 0;
        __syncthreads();  // staged values visible to the walker; buffer `cur` complete; counter armed
        if (has_next && warp == kWarps - 1) {
            if (COMPACT) {
                const unsigned long long stream = p.seed + nidx;
                for (int k = lane; k < d; k += 32) {
                    uint32_t ka, kb;
                    lhd2_column_key((uint32_t)k, (uint32_t)stream, (uint32_t)(stream >> 32), ka, kb);
                    fy_walk_column_v2(xbase + (cur ^ 1) * xstride + k * ns, ka, kb, n, p.c1, p.c0);
                }
            } else {
                cta_walk_columns(xbase + (cur ^ 1) * xstride, jst, n, d, ns, lane, 32);
            }
        }

        // ---- scoring: items (row tile, column block) in tile-major order from the shared counter
        const double* xs = xbase + cur * xstride;
        CtaScore st;
        st.reset();
        int t = 0, base = 0;
        int nblk = (n + kJB - 1) / kJB;  // column blocks from the tile's first row
        int grabbed = 0;
        if (lane == 0) grabbed = atomicAdd(&next_item, 1);
        while (true) {
            const int item = __shfl_sync(0xffffffffu, grabbed, 0);
            if (lane == 0) grabbed = atomicAdd(&next_item, 1);  // the next one, in flight while this one is scored
            while (t < tiles && item - base >= nblk) { base += nblk; ++t; nblk = (n - t * kTileRows + kJB - 1) / kJB; }
            if (t >= tiles) break;
            cta_score_item_any<MAXPRO>(chunk, xs, n, d, ns, t * kTileRows, t * kTileRows + (item - base) * kJB, lane, st);
        }
        if (MAXPRO) st.flush();
        const double part = MAXPRO ? warp_reduce_sum(st.acc) : warp_reduce_min(st.mn);
        if (lane == 0) red[warp] = part;
        __syncthreads();  // all partial sums in; the walk of the next candidate is done
        if (tid == 0) {
            double s = red[0];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s = MAXPRO ? s + red[w] : fmin(s, red[w]);
            if (!MAXPRO) s = sqrt(s);
            lb.offer(s, idx, p.n_pairs, p.inv_d);
        }
    }
    block_publish_best<MAXPRO>(lb.best, lb.idx, warp_best, p.block_best, p.done_counter, p.result);
}

static size_t cta_smem_bytes_layout(unsigned long long n, unsigned long long d, bool compact) {
    unsigned long long ns = cta_row_stride(n);
    return (size_t)(((d * ns * (compact ? 8 : 10) + 15) / 16) * 16 + 16 * sizeof(double));
}
// The compact layout (stream v2) takes over where the design no longer fits beside its side array -- and on wide designs,
// d >= 12, where it is a little faster: one lane per column walks, so d columns are generated side by side and the staging
// phase with its barrier is gone ((1000,20) 19.2 -> 20.1, (2000,10) 19.9 -> 20.2 algorithmic TFLOP/s; at d <= 10 the two
// layouts are within 1 %, tools/cta_variants.py).  The double-buffered kernel, which stages nothing either under stream v2,
// is taken whenever two of its CTAs fit one SM.
static bool cta_wide(unsigned long long d) { return stream_version() == 2 && d >= 12; }
static bool cta_compact(unsigned long long n, unsigned long long d) {
    if (stream_version() != 2) return false;
    if (const char* e = cfg("MAXPRO_CTA_COMPACT")) return atoi(e) != 0;
    return cta_wide(d) || cta_smem_bytes_layout(n, d, false) + 256 > kSmemBudget;
}
static size_t cta_smem_bytes(unsigned long long n, unsigned long long d) { return cta_smem_bytes_layout(n, d, cta_compact(n, d)); }

static bool cta_pipe_compact() {
    if (stream_version() != 2) return false;
    if (const char* e = cfg("MAXPRO_CTA_PIPE_COMPACT")) return atoi(e) != 0;
    return true;
}
static size_t cta_pipe_smem_bytes(unsigned long long n, unsigned long long d) {
    unsigned long long ns = cta_row_stride(n);
    return (size_t)(((d * ns * (cta_pipe_compact() ? 16 : 18) + 15) / 16) * 16);
}

// The pipelined kernel pays for its second buffer with occupancy; it is used while two CTAs of
// eight warps still fit one SM (the single-buffer kernel then runs four CTAs of four warps).
static bool cta_use_pipe(unsigned long long n, unsigned long long d) {
    if (const char* e = cfg("MAXPRO_CTA_PIPE")) return atoi(e) != 0 && cta_pipe_smem_bytes(n, d) + 1024 <= kSmemBudget;
    return 2 * (cta_pipe_smem_bytes(n, d) + 2048) <= kSmemBudget;
}

// CTAs of the single-buffer kernel per SM: 4 x 128 threads, 2 x 256 or 1 x 512 (MAXPRO_CTA_THREADS forces a block size)
static int cta_per_sm(unsigned long long n, unsigned long long d) {
    int per_sm = (int)(kSmemBudget / (cta_smem_bytes(n, d) + 1024));
    per_sm = per_sm >= 4 ? 4 : (per_sm >= 2 ? 2 : 1);
    if (const char* e = cfg("MAXPRO_CTA_THREADS")) {
        const int t = atoi(e);
        if ((t == 128 || t == 256 || t == 512) && 512 / t <= per_sm) per_sm = 512 / t;
    }
    return per_sm;
}

bool search_cta_supported(unsigned long long n, unsigned long long d) {
    return n >= 2 && n <= 65535 && d >= 1 && cta_smem_bytes(n, d) + 256 <= kSmemBudget;  // 256: static shared memory
}

int search_cta_grid(unsigned long long n, unsigned long long d, int sms) {
    if (cta_use_pipe(n, d)) {
        int per_sm = (int)(kSmemBudget / (cta_pipe_smem_bytes(n, d) + 2048));
        if (per_sm > 2) per_sm = 2;
        if (per_sm < 1) per_sm = 1;
        return sms * per_sm;
    }
    return sms * cta_per_sm(n, d);
}

cudaError_t launch_search_cta(const SearchLaunch& L) {
    CtaParams p;
    p.n = (int)L.n; p.d = (int)L.d; p.ns = (int)cta_row_stride(L.n);
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.gen = stream_version();
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = lhd_c0(p.gen, p.c1);
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.block_best = L.block_best; p.done_counter = L.done_counter; p.result = L.result;
    cudaError_t e;
    if (cta_use_pipe(L.n, L.d)) {
        const size_t psmem = cta_pipe_smem_bytes(L.n, L.d);
        e = cudaSuccess;
        auto gop = [&](auto kernel) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
            if (e != cudaSuccess) return;
            kernel<<<L.grid, kPipeThreads, psmem, L.stream>>>(p);
        };
        const bool pc = cta_pipe_compact();
        if (L.metric == 0) { if (pc) gop(search_cta_pipe_kernel<true, true>); else gop(search_cta_pipe_kernel<true, false>); }
        else { if (pc) gop(search_cta_pipe_kernel<false, true>); else gop(search_cta_pipe_kernel<false, false>); }
        if (e != cudaSuccess) return e;
        return cudaGetLastError();
    }
    const size_t smem = cta_smem_bytes(L.n, L.d);
    const int per_sm = cta_per_sm(L.n, L.d);
    e = cudaSuccess;
    auto go = [&](auto kernel, int threads) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        kernel<<<L.grid, threads, smem, L.stream>>>(p);
    };
    auto pick = [&](auto mp_tag, auto compact_tag) {
        constexpr bool MP = decltype(mp_tag)::value, CP = decltype(compact_tag)::value;
        if (per_sm >= 4) go(search_cta_kernel<MP, 128, CP>, 128);
        else if (per_sm == 2) go(search_cta_kernel<MP, 256, CP>, 256);
        else go(search_cta_kernel<MP, 512, CP>, 512);
    };
    const bool compact = cta_compact(L.n, L.d);
    if (L.metric == 0) { if (compact) pick(std::true_type{}, std::true_type{}); else pick(std::true_type{}, std::false_type{}); }
    else { if (compact) pick(std::false_type{}, std::true_type{}); else pick(std::false_type{}, std::false_type{}); }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace mp

// K1+K2 fused, "small design" shape: one THREAD per candidate.
//
// Replaces the map/reduce of src/lib.rs:65-98 for designs small enough that a
// CTA can hold one per thread in shared memory (n*d*8 B per thread, e.g. 1200 B
// at n=50, d=3): generate_lhd (src/lhd.rs:97-132) -> criterion
// (src/maxpro_utils.rs:28-83 or src/maximin_utils.rs:54-67) -> keep the best.
//
// Why thread-per-candidate: the i<j triangle of a 50-point design is 1225 pairs;
// spreading it over a warp wastes lanes on the triangle and needs shuffles for
// every reduction.  Here each lane owns a whole candidate, so all 32 lanes do
// useful FP64 work every cycle, control flow is uniform (same n, d for every
// lane), there is no __syncthreads and no shuffle in the hot loop, and warps
// drift apart so one warp's integer-only Philox phase overlaps another's FP64
// scoring phase.
//
// Shared-memory layout: X[k][i][t] (t = thread, fastest) -- element (k,i) of
// thread t lives at smem[(k*n + i)*T + t].  Any per-thread element index then
// maps lane t to bank pair (t mod 16): the data-dependent Fisher-Yates accesses
// and the tile loads are conflict-free by construction.
//
// FP64 budget per pair at d=3 (the roofline unit is 3d+3 = 12 algorithmic flops):
//   3 DADD (differences) + 2 DMUL (product of differences) + 1 DFMA (q*q + eps)
//   + 1 DFMA + 1 DMUL (RecipChain::push)                        = 8 pipe slots
//   + one reciprocal (MUFU + 5 DFMA) per <=16 pairs per chain   ~ 0.4
// against ~15 for the literal sub/sq/mul/add/divide body.
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cstdlib>

namespace mp {

struct SearchParams {
    int n;
    int gen;              // design stream version (philox.cuh)
    unsigned long long seed, lo, hi;
    double c1, c0;        // 2^-32/n and 2^-33/n : v = (i + (jit+0.5)*2^-32)/n
    double n_pairs, inv_d;
    // list mode (stage 2 of the two-stage search, search_screen.cu): score list[0..*list_count_dev) instead of lo..hi;
    // a count above list_cap means stage 1 overflowed its list: the whole range is scored instead
    const unsigned long long* list;
    const unsigned int* list_count_dev;
    unsigned int list_cap;
    BestRec* block_best;  // [gridDim.x]
    unsigned int* done_counter;
    BestRec* result;      // final {score, index}
};

template <int D>
__device__ __forceinline__ double maxpro_term(const double (&a)[D], const double (&b)[D]) {
    double q = a[0] - b[0];
#pragma unroll
    for (int k = 1; k < D; ++k) q *= (a[k] - b[k]);
    return fma(q, q, kEps);  // (prod diff)^2 + eps == prod diff^2 + eps up to 1 ulp
}

// Bit-exact restatement of calculate_l2_distance before the sqrt
// (src/maximin_utils.rs:39-44): separate multiply and add, left to right from 0.
template <int D>
__device__ __forceinline__ double sqdist_term(const double (&a)[D], const double (&b)[D]) {
    double df = __dsub_rn(a[0], b[0]);
    double s = __dmul_rn(df, df);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        df = __dsub_rn(a[k], b[k]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

// One column of generate_lhd: inside-out Fisher-Yates over the strata, the
// in-stratum jitter attached to the value being placed.
__device__ __forceinline__ void generate_column(int gen, double* col, int stride, int n, uint32_t k,
                                                uint32_t s_lo, uint32_t s_hi, double c1, double c0) {
    LhdKeyCache kc{0xFFFFFFFFu, 0u, 0u};
    for (int i = 0; i < n; i += 2) {
        const LhdBlock w = lhd_draw_block(gen, (uint32_t)(i >> 1), k, s_lo, s_hi, kc);
        {
            const double v = lhd_value(gen, (uint32_t)i, w.jit0, c1, c0);
            col[i * stride] = col[w.j0 * stride];
            col[w.j0 * stride] = v;
        }
        if (i + 1 < n) {
            const double v = lhd_value(gen, (uint32_t)i + 1u, w.jit1, c1, c0);
            col[(i + 1) * stride] = col[w.j1 * stride];
            col[w.j1 * stride] = v;
        }
    }
}

constexpr int kMaxChain = kMaxChainTerms;  // terms per RecipChain (started at 2^1000, common.cuh) before a flush

// Score one design held in shared memory by a group of G lanes (G = 1 or 2).
// MAXPRO: returns this lane's share of S = sum_{i<j} 1/(prod diff^2 + eps)
// (src/maxpro_utils.rs:38-58); !MAXPRO: this lane's min_{i<j} squared distance.
// Work split (uniform control flow, so lanes never diverge):
//   phase A  every lane loads the same R-row tile into registers and the rows after
//            the tile are dealt round-robin to the G lanes (R pairs per loaded row);
//   phase B  the R(R-1)/2 pairs inside a tile are taken tile-by-tile, tiles dealt
//            round-robin to the lanes;
//   phase C  pairs among the < R rows left over when R does not divide n.
template <int D, int R, int G, bool MAXPRO>
__device__ __forceinline__ double score_group_design(const double* xs, int stride, int n, int g) {
    constexpr int U = (D >= 5) ? 2 : ((R == 5) ? 5 : 4);  // rows per unrolled body (register budget)
    const int kstride = n * stride;      // distance between dimensions
    RecipChain ch[R];
    double mins[R];
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { ch[r].reset(); mins[r] = CUDART_INF; }
    int pushed = 0;  // uniform upper bound on terms held by any chain

    auto flush_all = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) acc = ch[r].flush_into(acc);
        pushed = 0;
    };
    auto load_row = [&](double (&dst)[D], int row) {
#pragma unroll
        for (int k = 0; k < D; ++k) dst[k] = xs[k * kstride + row * stride];
    };

    const int full = (n / R) * R;
    // ---- phase A
    for (int i0 = 0; i0 < full; i0 += R) {
        double xi[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r) load_row(xi[r], i0 + r);
        int jb = i0 + R;
        for (; jb + G * U <= n; jb += G * U) {
            if (MAXPRO && pushed + U > kMaxChain) flush_all();
            double xj[U][D];
#pragma unroll
            for (int u = 0; u < U; ++u) load_row(xj[u], jb + g + G * u);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (MAXPRO) ch[r].push(maxpro_term<D>(xi[r], xj[u]));
                    else mins[r] = min_sqdist(mins[r], sqdist_term<D>(xi[r], xj[u]));
                }
            if (MAXPRO) pushed += U;
        }
        for (; jb < n; jb += G) {
            if (MAXPRO && pushed + 1 > kMaxChain) flush_all();
            const int j = jb + g;
            if (G == 1 || j < n) {
                double xj[D];
                load_row(xj, j);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (MAXPRO) ch[r].push(maxpro_term<D>(xi[r], xj));
                    else mins[r] = min_sqdist(mins[r], sqdist_term<D>(xi[r], xj));
                }
            }
            if (MAXPRO) pushed += 1;
        }
    }
    // ---- phase B
    for (int i0 = g * R; i0 < full; i0 += G * R) {
        double xi[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r) load_row(xi[r], i0 + r);
        if (MAXPRO && pushed + R > kMaxChain) flush_all();
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int r2 = r + 1; r2 < R; ++r2) {
                if (MAXPRO) ch[(r + r2) % R].push(maxpro_term<D>(xi[r], xi[r2]));
                else mins[(r + r2) % R] = min_sqdist(mins[(r + r2) % R], sqdist_term<D>(xi[r], xi[r2]));
            }
        if (MAXPRO) pushed += R / 2 + 1;
    }
    // ---- phase C
    int turn = 0;
    for (int i = full; i < n; ++i) {
        double xi[D];
        load_row(xi, i);
        for (int j = i + 1; j < n; ++j, ++turn) {
            if (MAXPRO && pushed + 1 > kMaxChain) flush_all();
            if (G == 1 || (turn % G) == g) {
                double xj[D];
                load_row(xj, j);
                if (MAXPRO) ch[0].push(maxpro_term<D>(xi, xj));
                else mins[0] = min_sqdist(mins[0], sqdist_term<D>(xi, xj));
            }
            if (MAXPRO) pushed += 1;
        }
    }
    if (MAXPRO) {
        flush_all();
        return acc;
    }
    double m = mins[0];
#pragma unroll
    for (int r = 1; r < R; ++r) m = min_sqdist(m, mins[r]);
    return m;
}

// Named-barrier helpers: a producer/consumer token between two warps (64 threads).
__device__ __forceinline__ void token_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void token_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// G lanes per candidate: lane g generates columns g, g+G, ... and scores its share of
// the pairs.  G = 2 doubles the resident warps for the same shared memory (12 warps per
// SM at n=50, d=3); G = 4 serves the designs of 2-3.5 KB with four or more columns, which
// otherwise run on four to six warps per SM -- one per sub-partition, nothing to overlap its
// generation phase with.
//
// LOCK: the warps of one SM sub-partition (warp id mod 4) take turns on the FP64 pipe.
// Every warp alternates an integer-only, latency-bound phase (Philox + Fisher-Yates) and an
// FP64-bound phase (the pair loop).  Left alone, the warps of a sub-partition share the pipe
// fairly, therefore finish scoring together, therefore all generate together while the FP64
// pipe idles -- measured 56% pipe utilisation.  Passing a token round the 2-3 warps of each
// sub-partition (named barriers, bar.arrive / bar.sync on 64 threads) makes the scoring phases
// strictly consecutive, so one warp's generation always overlaps another warp's scoring.
template <int D, int R, int G, bool MAXPRO, bool LOCK>
__global__ void __launch_bounds__(512, 1) search_small_kernel(SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = blockDim.x / G;  // candidates resident in this CTA
    const int n = p.n;
    // Lane layout inside a warp: lanes [0, 32/G) are lane-role 0 of 32/G consecutive
    // candidates, lanes [32/G, 2*32/G) lane-role 1 of the same candidates, and so on.  With
    // G <= 2 each half-warp (the unit a 64-bit shared-memory access is split into) touches 16
    // different candidates = 16 different bank pairs: no conflicts for any per-lane element
    // index.  With G = 4 a half-warp holds two roles of 8 candidates; the host then picks C = 8
    // (mod 16), which puts rows of different parity into different halves of the banks: the
    // partner rows of phase A (roles differ by one row) stay conflict-free, the same row read by
    // both roles is a broadcast, and only the data-dependent Fisher-Yates accesses collide (half
    // of the time, two-way).
    constexpr int kPerWarp = 32 / G;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = wid * kPerWarp + (lane % kPerWarp), g = lane / kPerWarp;
    unsigned int group_mask = 0u;  // the G lanes of this candidate
#pragma unroll
    for (int j = 0; j < G; ++j) group_mask |= 1u << (lane % kPerWarp + j * kPerWarp);
    double* xs = reinterpret_cast<double*>(smem_raw) + c;
    BestRec* warp_best = reinterpret_cast<BestRec*>(smem_raw + (size_t)n * D * C * sizeof(double));

    const unsigned long long gcid = (unsigned long long)blockIdx.x * C + c;
    const unsigned long long gstep = (unsigned long long)gridDim.x * C;
    const unsigned long long* list = p.list;
    unsigned long long list_count = 0;
    if (list) {
        list_count = *p.list_count_dev;
        if (list_count > p.list_cap) list = nullptr;
    }
    const unsigned long long total = list ? list_count : p.hi - p.lo;
    const unsigned long long rounds = (total + gstep - 1) / gstep;  // same trip count for every thread

    // token ring of the warps sharing this warp's sub-partition: barrier 1 + smsp*ring + q is
    // the token INTO ring position q
    const int ring = (blockDim.x >> 5) >> 2, smsp = wid & 3, q = wid >> 2;
    const int tok_in = 1 + smsp * ring + q, tok_out = 1 + smsp * ring + (q + 1 == ring ? 0 : q + 1);
    if (LOCK && q == ring - 1) token_pass(tok_out);  // position 0 scores first

    LaneBest<MAXPRO> lb;  // fold of src/lib.rs:85-97 on the criterion (reduce.cuh)
    lb.init(p.inv_d);

    for (unsigned long long rnd = 0; rnd < rounds; ++rnd) {
        const unsigned long long off = gcid + rnd * gstep;
        const bool active = off < total;
        const unsigned long long idx = list ? (active ? list[off] : 0ull) : p.lo + off;
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        if (active) {
#pragma unroll 1
            for (int k = g; k < D; k += G)
                generate_column(p.gen, xs + (size_t)k * n * C, C, n, (uint32_t)k, s_lo, s_hi, p.c1, p.c0);
        }
        if (G > 1) __syncwarp(group_mask);
        if (LOCK) token_wait(tok_in);
        double s = 0.0;
        if (active) s = score_group_design<D, R, G, MAXPRO>(xs, C, n, g);
        if (LOCK) token_pass(tok_out);
        if (G > 1) {
            // fixed order (lower role + higher role at every level) so all lanes of the group hold the same bits
#pragma unroll
            for (int step = 1; step < G; step <<= 1) {
                const double other = __shfl_xor_sync(group_mask, s, step * kPerWarp);
                if (MAXPRO) s = ((g & step) == 0) ? s + other : other + s;
                else s = fmin(s, other);
            }
            __syncwarp(group_mask);  // partners are done reading before the design is overwritten
        }
        if (!MAXPRO) s = sqrt(s);  // reference compares distances, not squares (ties)
        if (active) lb.offer(s, idx, p.n_pairs, p.inv_d);
    }
    if (LOCK && q == 0) token_wait(tok_in);  // consume the last token so no arrival is left pending
    block_publish_best<MAXPRO>(lb.best, lb.idx, warp_best, p.block_best, p.done_counter, p.result);
}

// ---- host side -------------------------------------------------------------

template <int D, int R, int G, bool MAXPRO, bool LOCK>
static cudaError_t launch_one(const SearchLaunch& L, const SearchParams& p, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(search_small_kernel<D, R, G, MAXPRO, LOCK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    search_small_kernel<D, R, G, MAXPRO, LOCK><<<L.grid, L.threads, smem, L.stream>>>(p);
    return cudaGetLastError();
}

template <int D, int R, int G>
static cudaError_t launch_drg(const SearchLaunch& L, const SearchParams& p) {
    size_t smem = (size_t)p.n * D * (L.threads / G) * sizeof(double) + 32 * sizeof(BestRec);
    // the token ring needs whole rings: warps a multiple of 4, 2 or 3 warps per sub-partition
    const int warps = L.threads / 32;
    bool lock = (L.threads % 128 == 0) && (warps / 4 == 2 || warps / 4 == 3);
    if (const char* e = cfg("MAXPRO_SMALL_LOCK")) lock = lock && atoi(e) != 0;
    if (L.metric == 0) return lock ? launch_one<D, R, G, true, true>(L, p, smem) : launch_one<D, R, G, true, false>(L, p, smem);
    return lock ? launch_one<D, R, G, false, true>(L, p, smem) : launch_one<D, R, G, false, false>(L, p, smem);
}

// Register tile height R by dimension count: the tile (R*D doubles) plus the unrolled rows
// must stay inside the 128-register budget of a 512-thread CTA.
template <int D>
static cudaError_t launch_d(const SearchLaunch& L, const SearchParams& p) {
    if constexpr (D <= 4) {
        const bool r5 = (p.n % 5 == 0);
        if constexpr (D == 4) { if (L.group == 4) return r5 ? launch_drg<D, 5, 4>(L, p) : launch_drg<D, 4, 4>(L, p); }
        if (L.group == 2) return r5 ? launch_drg<D, 5, 2>(L, p) : launch_drg<D, 4, 2>(L, p);
        return r5 ? launch_drg<D, 5, 1>(L, p) : launch_drg<D, 4, 1>(L, p);
    } else if constexpr (D <= 6) {
        if (L.group == 4) return launch_drg<D, 4, 4>(L, p);
        return L.group == 2 ? launch_drg<D, 4, 2>(L, p) : launch_drg<D, 4, 1>(L, p);
    } else {
        if (L.group == 4) return launch_drg<D, 3, 4>(L, p);
        return L.group == 2 ? launch_drg<D, 3, 2>(L, p) : launch_drg<D, 3, 1>(L, p);
    }
}

// Shapes this kernel takes: at least 64 candidates must fit one SM's shared memory (n*d*8 B <= ~3.5 KB).
// Measured against the warp-per-candidate kernel (tools/shape_sweep.py, after that kernel moved to
// row-pair items): at 3.0-3.2 KB this kernel leads ((100,4) 13.3 vs 10.1, (80,5) 11.5 vs 9.9, (50,8)
// 8.7 vs 6.2 algorithmic TFLOP/s), from 3.8 KB up the warp kernel does ((100,5) 9.3 vs 10.7, (300,2)
// 12.7 vs 14.0, (75,8) 7.9 vs 9.6).
bool search_small_supported(unsigned long long n, unsigned long long d) {
    unsigned long long min_cands = 64;
    if (const char* e = cfg("MAXPRO_SMALL_MIN_CANDS")) { int v = atoi(e); if (v >= 16 && v <= 512 && v % 16 == 0) min_cands = (unsigned long long)v; }
    // 256 bytes of slack: the kernel's static shared memory (block_publish_best) and the per-CTA reservation come out of the
    // same 227 KB -- a design that fills the budget to the byte, (151,3), failed its launch with cudaErrorInvalidValue
    return d >= 1 && d <= kSmallMaxD && n >= 2 && n * d * sizeof(double) * min_cands + 32 * sizeof(BestRec) + 256 <= kSmemBudget;
}

// Lanes per candidate and CTA size.  Two lanes per candidate whenever there are at least two
// columns to generate; as many candidates as fit one SM's shared memory (multiple of 16 so
// every lane of a half-warp hits its own bank pair), at most 512 threads.
void search_small_pick_config(unsigned long long n, unsigned long long d, int* group, int* threads) {
    int G = d >= 2 ? 2 : 1;
    size_t per = (size_t)n * d * sizeof(double);
    size_t cap = (kSmemBudget - 32 * sizeof(BestRec) - 256) / per;  // 256: the kernel's static shared memory (see search_small_supported)
    // Four lanes per candidate when shared memory leaves two lanes fewer than 10 warps (n*d above ~180
    // elements) and there are four or more columns -- except d = 5 (2+1+1+1 columns per lane: the
    // generation phase is as long as with two lanes, and the scoring tiles get smaller).  Measured,
    // G = 2 -> 4 in algorithmic TFLOP/s: (50,4) 11.9 -> 13.9, (50,6) 10.5 -> 11.9, (50,8) 8.7 -> 10.1,
    // (40,8) 6.8 -> 10.9, (100,4) 12.6 -> 13.8, (25,8) 7.6 -> 9.0, (60,7) 9.1 -> 10.9; against it
    // (50,5) 11.6 -> 10.3, and the designs that already fill 10+ warps: (30,6) 10.2 -> 9.5, (40,4)
    // 13.1 -> 12.5, (20,4) 10.0 -> 7.6.
    if (d >= 4 && d != 5 && (cap / 16) * 16 * 2 < 320) G = 4;
    if (const char* e = cfg("MAXPRO_SMALL_GROUP")) { int v = atoi(e); if (v == 1 || v == 2 || (v == 4 && d >= 4)) G = v; }
    size_t c;
    if (G == 4) {
        c = cap >= 8 ? ((cap - 8) / 16) * 16 + 8 : 0;  // 8 (mod 16): see the lane layout note in the kernel
        if (c > 120) c = 120;                           // 480 threads
        if (c == 0) { G = 2; }
    }
    if (G != 4) {
        c = (cap / 16) * 16;
        if (c * G > 512) c = 512 / G;
        if (const char* e = cfg("MAXPRO_SMALL_CANDS")) { size_t v = (size_t)atoi(e); if (v >= 16 && v <= c && v % 16 == 0) c = v; }
        if (G == 1) c = (c / 32) * 32;
    }
    *group = G;
    *threads = (int)(c * G);
}

cudaError_t launch_search_small(const SearchLaunch& L) {
    SearchParams p;
    p.n = (int)L.n;
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.gen = stream_version();
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = lhd_c0(p.gen, p.c1);
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.list = L.list; p.list_count_dev = L.list_count_dev; p.list_cap = L.list_cap;
    p.block_best = L.block_best; p.done_counter = L.done_counter; p.result = L.result;
    switch (L.d) {
        case 1: return launch_d<1>(L, p);
        case 2: return launch_d<2>(L, p);
        case 3: return launch_d<3>(L, p);
        case 4: return launch_d<4>(L, p);
        case 5: return launch_d<5>(L, p);
        case 6: return launch_d<6>(L, p);
        case 7: return launch_d<7>(L, p);
        case 8: return launch_d<8>(L, p);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mp

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

namespace mp {

struct SearchParams {
    int n;
    unsigned long long seed, lo, hi;
    double c1, c0;        // 2^-32/n and 2^-33/n : v = (i + (jit+0.5)*2^-32)/n
    double n_pairs, inv_d;
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
__device__ __forceinline__ void generate_column(double* col, int stride, int n, uint32_t k,
                                                uint32_t s_lo, uint32_t s_hi, double c1, double c0) {
    for (int i = 0; i < n; i += 2) {
        Philox4 w = philox4x32_10((uint32_t)(i >> 1), k, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
        {
            uint32_t j = mulhi_bound(w.y, (uint32_t)i + 1u);
            unsigned long long m = ((unsigned long long)(uint32_t)i << 32) | w.x;
            double v = fma((double)m, c1, c0);
            col[i * stride] = col[j * stride];
            col[j * stride] = v;
        }
        if (i + 1 < n) {
            uint32_t j = mulhi_bound(w.w, (uint32_t)i + 2u);
            unsigned long long m = ((unsigned long long)(uint32_t)(i + 1) << 32) | w.z;
            double v = fma((double)m, c1, c0);
            col[(i + 1) * stride] = col[j * stride];
            col[j * stride] = v;
        }
    }
}

constexpr int kMaxChain = 16;  // terms per RecipChain before a flush (den >= 1e-192)
constexpr int kJUnroll = 4;    // j iterations per unrolled body

// Score the thread's design.  MAXPRO: returns S = sum_{i<j} 1/(prod diff^2 + eps)
// (src/maxpro_utils.rs:38-58).  !MAXPRO: returns min_{i<j} squared distance.
template <int D, int R, bool MAXPRO>
__device__ __forceinline__ double score_thread_design(const double* xs, int stride, int n) {
    const int kstride = n * stride;  // distance between dimensions
    RecipChain ch[R];
    double mins[R];
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { ch[r].reset(); mins[r] = CUDART_INF; }
    int pushed = 0;  // uniform upper bound on terms held by any chain

    auto flush_all = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) acc += ch[r].flush();
        pushed = 0;
    };

    int i0 = 0;
    for (; i0 + R <= n; i0 += R) {
        double xi[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < D; ++k) xi[r][k] = xs[k * kstride + (i0 + r) * stride];

        // pairs inside the row tile, spread over the chains
        if (MAXPRO && pushed + R > kMaxChain) flush_all();
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int r2 = r + 1; r2 < R; ++r2) {
                if (MAXPRO) ch[(r + r2) % R].push(maxpro_term<D>(xi[r], xi[r2]));
                else mins[(r + r2) % R] = fmin(mins[(r + r2) % R], sqdist_term<D>(xi[r], xi[r2]));
            }
        if (MAXPRO) pushed += R / 2 + 1;

        // the tile against every later row: R pairs per loaded row
        int j = i0 + R;
        for (; j + kJUnroll <= n; j += kJUnroll) {
            if (MAXPRO && pushed + kJUnroll > kMaxChain) flush_all();
            double xj[kJUnroll][D];
#pragma unroll
            for (int u = 0; u < kJUnroll; ++u)
#pragma unroll
                for (int k = 0; k < D; ++k) xj[u][k] = xs[k * kstride + (j + u) * stride];
#pragma unroll
            for (int u = 0; u < kJUnroll; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (MAXPRO) ch[r].push(maxpro_term<D>(xi[r], xj[u]));
                    else mins[r] = fmin(mins[r], sqdist_term<D>(xi[r], xj[u]));
                }
            if (MAXPRO) pushed += kJUnroll;
        }
        for (; j < n; ++j) {
            if (MAXPRO && pushed + 1 > kMaxChain) flush_all();
            double xj[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xj[k] = xs[k * kstride + j * stride];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (MAXPRO) ch[r].push(maxpro_term<D>(xi[r], xj));
                else mins[r] = fmin(mins[r], sqdist_term<D>(xi[r], xj));
            }
            if (MAXPRO) pushed += 1;
        }
    }
    // rows left over when R does not divide n
    for (int i = i0; i < n; ++i) {
        double xi[D];
#pragma unroll
        for (int k = 0; k < D; ++k) xi[k] = xs[k * kstride + i * stride];
        for (int j = i + 1; j < n; ++j) {
            if (MAXPRO && pushed + 1 > kMaxChain) flush_all();
            double xj[D];
#pragma unroll
            for (int k = 0; k < D; ++k) xj[k] = xs[k * kstride + j * stride];
            if (MAXPRO) { ch[0].push(maxpro_term<D>(xi, xj)); pushed += 1; }
            else mins[0] = fmin(mins[0], sqdist_term<D>(xi, xj));
        }
    }
    if (MAXPRO) {
        flush_all();
        return acc;
    }
    double m = mins[0];
#pragma unroll
    for (int r = 1; r < R; ++r) m = fmin(m, mins[r]);
    return m;
}

template <int D, int R, bool MAXPRO>
__global__ void __launch_bounds__(256, 1) search_small_kernel(SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int n = p.n;
    double* xs = reinterpret_cast<double*>(smem_raw) + threadIdx.x;
    BestRec* warp_best = reinterpret_cast<BestRec*>(smem_raw + (size_t)n * D * T * sizeof(double));

    const unsigned long long gtid = (unsigned long long)blockIdx.x * T + threadIdx.x;
    const unsigned long long gsize = (unsigned long long)gridDim.x * T;

    // MAXPRO minimises S (psi is monotone in S); maximin maximises the distance.
    double best = MAXPRO ? CUDART_INF : -CUDART_INF;
    unsigned long long best_idx = kNoIndex;

    for (unsigned long long idx = p.lo + gtid; idx < p.hi; idx += gsize) {
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
#pragma unroll 1
        for (int k = 0; k < D; ++k)
            generate_column(xs + (size_t)k * n * T, T, n, (uint32_t)k, s_lo, s_hi, p.c1, p.c0);
        double s = score_thread_design<D, R, MAXPRO>(xs, T, n);
        if (!MAXPRO) s = sqrt(s);  // reference compares distances, not squares (ties)
        // idx grows along the loop, so ">= / <=" keeps the later index on ties
        if (MAXPRO ? (s <= best) : (s >= best)) { best = s; best_idx = idx; }
    }
    if (MAXPRO && best_idx != kNoIndex) best = pow(best / p.n_pairs, p.inv_d);  // maxpro_utils.rs:81-82
    block_publish_best<MAXPRO>(best, best_idx, warp_best, p.block_best, p.done_counter, p.result);
}

// ---- host side -------------------------------------------------------------

template <int D, int R>
static cudaError_t launch_dr(const SearchLaunch& L, const SearchParams& p) {
    size_t smem = (size_t)p.n * D * L.threads * sizeof(double) + 32 * sizeof(BestRec);
    cudaError_t e;
    if (L.metric == 0) {
        e = cudaFuncSetAttribute(search_small_kernel<D, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_small_kernel<D, R, true><<<L.grid, L.threads, smem, L.stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(search_small_kernel<D, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_small_kernel<D, R, false><<<L.grid, L.threads, smem, L.stream>>>(p);
    }
    return cudaGetLastError();
}

template <int D>
static cudaError_t launch_d(const SearchLaunch& L, const SearchParams& p) {
    if (p.n % 5 == 0 && D <= 4) return launch_dr<D, 5>(L, p);
    return launch_dr<D, 4>(L, p);
}

bool search_small_supported(unsigned long long n, unsigned long long d) {
    return d >= 1 && d <= kSmallMaxD && n >= 2 && n * d * sizeof(double) * 64 + 32 * sizeof(BestRec) <= kSmemBudget;
}

int search_small_pick_threads(unsigned long long n, unsigned long long d) {
    // Largest warp multiple (<= 256) whose designs fit in one SM's shared memory.
    size_t per_thread = (size_t)n * d * sizeof(double);
    size_t t = (kSmemBudget - 32 * sizeof(BestRec)) / per_thread;
    t = (t / 32) * 32;
    if (t > 256) t = 256;
    return (int)t;
}

cudaError_t launch_search_small(const SearchLaunch& L) {
    SearchParams p;
    p.n = (int)L.n;
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = 0.5 * p.c1;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
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

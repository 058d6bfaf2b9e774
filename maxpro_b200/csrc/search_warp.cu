// K1+K2 fused, "medium design" shape: one WARP per candidate (e.g. n=100, d=10: 10 KB).
//
// Replaces the same reference code as the other search kernels: generate_lhd
// (src/lhd.rs:97-132) -> criterion (src/maxpro_utils.rs:28-83 / src/maximin_utils.rs:54-67)
// -> fold (src/lib.rs:76-98).  It fills the gap between the thread-group kernels
// (search_small.cu: designs up to 4.8 KB) and the CTA-per-candidate kernel (search_cta.cu):
// at n ~ 100-200 a CTA spends most of a candidate in the sequential Fisher-Yates walk (one
// lane per column, n dependent steps) and in half-empty diagonal tiles, and its 128 registers
// x 128 threads allow only 4 candidates in flight per SM.  Here 16 warps per SM each carry
// their own candidate in shared memory, so the walks of some overlap the scoring of others.
//
// Generation (LHD-Philox v1, bit-identical to generate.cu / the oracle): as in search_cta.cu --
// every lane runs Philox for its share of the blocks and stages v_i and the Fisher-Yates
// target j_i; then one lane per column walks the n dependent swaps.
// Scoring: pairs enumerated cyclically, {i, (i+s) mod n} for s = 1..(n-1)/2 (+ the n/2 diameters
// of the lower half for even n), as work items of (row i, 8 consecutive shifts).  The items
// are dealt to the lanes in row-major order (consecutive lanes <-> consecutive rows: conflict-
// free), so every lane gets the same number of items whatever n is.  The 8 running products
// stay in registers across the dimensions; MaxPro terms go through the reciprocal chains of
// common.cuh (candidates are in the unit cube by construction).  Each column is followed by a
// copy of its first 8*ceil(smax/8) rows, so row (i+s) mod n is simply row i+s: the 8 partner
// loads of an item are one address plus immediates, with no wrap arithmetic and no index array.
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cstdlib>

namespace mp {

constexpr int kWarpCtaThreads = 128;  // 4 candidates per CTA, up to 4 CTAs per SM
constexpr int kWarpSB = 8;            // shifts per work item

struct WarpParams {
    int n, d, ns;  // ns = odd row stride of the [k][i] layout (the column walkers hit different banks): n + copied rows
    unsigned int cand_bytes;  // shared memory of one candidate slot
    unsigned long long seed, lo, hi;
    double c1, c0, n_pairs, inv_d;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

template <bool MAXPRO>
__global__ void __launch_bounds__(kWarpCtaThreads, 4) search_warp_kernel(WarpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BestRec warp_best[kWarpCtaThreads / 32];
    const int n = p.n, d = p.d, ns = p.ns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = reinterpret_cast<double*>(smem_raw + (size_t)warp * p.cand_bytes);  // [d][ns]
    unsigned short* jst = reinterpret_cast<unsigned short*>(xs + (size_t)d * ns);    // [d][n]
    const int nb = (n + 1) / 2;
    const int half = (n - 1) / 2;
    const bool even = (n & 1) == 0;
    const int nblk = (half + (even ? 1 : 0) + kWarpSB - 1) / kWarpSB;  // shift blocks per row
    const int items = n * nblk;

    double best = MAXPRO ? CUDART_INF : -CUDART_INF;  // kept by lane 0 of each warp
    unsigned long long best_idx = kNoIndex;

    const unsigned long long wstep = (unsigned long long)gridDim.x * (kWarpCtaThreads / 32);
    for (unsigned long long idx = p.lo + (unsigned long long)blockIdx.x * (kWarpCtaThreads / 32) + warp; idx < p.hi; idx += wstep) {
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        __syncwarp();  // previous candidate fully consumed
        // ---- generation, phase 1
        for (int q = lane; q < d * nb; q += 32) {
            const int k = q / nb, b = q - k * nb;
            Philox4 w = philox4x32_10((uint32_t)b, (uint32_t)k, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
            const int i = 2 * b;
            xs[k * ns + i] = stratum_value((uint32_t)i, w.x, p.c1, p.c0);
            jst[k * n + i] = (unsigned short)mulhi_bound(w.y, (uint32_t)i + 1u);
            if (i + 1 < n) {
                xs[k * ns + i + 1] = stratum_value((uint32_t)(i + 1), w.z, p.c1, p.c0);
                jst[k * n + i + 1] = (unsigned short)mulhi_bound(w.w, (uint32_t)i + 2u);
            }
        }
        __syncwarp();
        // ---- generation, phase 2: one lane per column
        for (int k = lane; k < d; k += 32) fy_walk_column(xs + k * ns, jst + k * n, n);
        __syncwarp();
        // rows n .. n + 8*nblk - 1 repeat rows 0 .. (wrapping again if the design is tiny)
        for (int e = lane; e < d * kWarpSB * nblk; e += 32) {
            const int k = e / (kWarpSB * nblk), r = e - k * (kWarpSB * nblk);
            xs[k * ns + n + r] = xs[k * ns + r % n];
        }
        __syncwarp();

        // ---- scoring: item q <-> (row q mod n, shift block q / n)
        double acc = 0.0, mn = CUDART_INF;
        RecipChain ch[kWarpSB];
#pragma unroll
        for (int c = 0; c < kWarpSB; ++c) ch[c].reset();
        int pushed = 0;
        int i = lane % n, b = lane / n;  // n may be below 32
        for (int q = lane; q < items; q += 32) {
            const int s0 = 1 + b * kWarpSB;
            const int my_smax = half + ((even && i < n / 2) ? 1 : 0);
            const double* pi = xs + i;        // row i
            const double* pj = pi + s0;       // rows i+s0 .. i+s0+7 (copies past row n-1)
            double pr[kWarpSB];
            {
                const double xi = pi[0];
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) pr[c] = xi - pj[c];
                    else { const double df = __dsub_rn(xi, pj[c]); pr[c] = __dmul_rn(df, df); }
                }
            }
            for (int k = 1; k < d; ++k) {
                pi += ns; pj += ns;
                const double xi = pi[0];
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) pr[c] *= xi - pj[c];
                    else { const double df = __dsub_rn(xi, pj[c]); pr[c] = __dadd_rn(pr[c], __dmul_rn(df, df)); }
                }
            }
            const bool whole = s0 + kWarpSB - 1 <= my_smax;  // false only in a row's last shift block
            if (MAXPRO) {
                if (pushed == kMaxChainTerms) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) acc = ch[c].flush_into(acc);
                    pushed = 0;
                }
                if (whole) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) ch[c].push(fma(pr[c], pr[c], kEps));
                } else {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c)
                        if (s0 + c <= my_smax) ch[c].push(fma(pr[c], pr[c], kEps));
                }
                ++pushed;
            } else {
                if (whole) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) mn = fmin(mn, pr[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c)
                        if (s0 + c <= my_smax) mn = fmin(mn, pr[c]);
                }
            }
            // next item of this lane: 32 rows further, wrapping into the next shift block
            i += 32;
            while (i >= n) { i -= n; ++b; }
        }
        if (MAXPRO) {
#pragma unroll
            for (int c = 0; c < kWarpSB; ++c) acc = ch[c].flush_into(acc);
        }
        double s = MAXPRO ? warp_reduce_sum(acc) : warp_reduce_min(mn);
        if (!MAXPRO) s = sqrt(s);
        if (MAXPRO ? (s <= best) : (s >= best)) { best = s; best_idx = idx; }  // later index on ties
    }
    if (MAXPRO && best_idx != kNoIndex) best = pow(best / p.n_pairs, p.inv_d);  // maxpro_utils.rs:81-82
    if (lane != 0) { best = MAXPRO ? CUDART_INF : -CUDART_INF; best_idx = kNoIndex; }
    block_publish_best<MAXPRO>(best, best_idx, warp_best, p.block_best, p.done_counter, p.result);
}

static unsigned long long warp_row_stride(unsigned long long n) {
    const unsigned long long smax = (n - 1) / 2 + ((n & 1) == 0 ? 1 : 0);
    const unsigned long long nblk = (smax + kWarpSB - 1) / kWarpSB;
    return (n + kWarpSB * nblk) | 1ull;  // n rows + the copied ones, odd
}
static unsigned int warp_cand_bytes(unsigned long long n, unsigned long long d) {
    return (unsigned int)(((d * warp_row_stride(n) * 8 + d * n * 2 + 15) / 16) * 16);
}

static int warp_ctas_per_sm(unsigned long long n, unsigned long long d) {
    const size_t per_cta = (size_t)warp_cand_bytes(n, d) * (kWarpCtaThreads / 32) + 1024;  // + static and reserved shared memory
    size_t c = kSmemBudget / per_cta;
    return (int)(c > 4 ? 4 : c);
}

// Medium designs: too big for the thread-group kernels, small enough that at least 12 candidates
// (3 CTAs of 4 warps) fit the shared memory of one SM -- below that the Fisher-Yates walks are no
// longer hidden and the CTA-per-candidate kernel does better.
bool search_warp_supported(unsigned long long n, unsigned long long d) {
    if (const char* e = getenv("MAXPRO_NO_WARP")) if (atoi(e) != 0) return false;
    int need = 3;
    if (const char* e = getenv("MAXPRO_WARP_MIN_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 4) need = v; }
    return n >= 2 && n <= 65535 && d >= 1 && d <= 65535 && warp_ctas_per_sm(n, d) >= need;
}

int search_warp_grid(unsigned long long n, unsigned long long d, int sms) { return warp_ctas_per_sm(n, d) * sms; }

cudaError_t launch_search_warp(const SearchLaunch& L) {
    WarpParams p;
    p.n = (int)L.n; p.d = (int)L.d; p.ns = (int)warp_row_stride(L.n);
    p.cand_bytes = warp_cand_bytes(L.n, L.d);
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = 0.5 * p.c1;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.block_best = L.block_best; p.done_counter = L.done_counter; p.result = L.result;
    const size_t smem = (size_t)p.cand_bytes * (kWarpCtaThreads / 32);
    cudaError_t e;
    if (L.metric == 0) {
        e = cudaFuncSetAttribute(search_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_warp_kernel<true><<<L.grid, kWarpCtaThreads, smem, L.stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(search_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_warp_kernel<false><<<L.grid, kWarpCtaThreads, smem, L.stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace mp

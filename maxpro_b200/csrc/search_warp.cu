// K1+K2 fused, "medium design" shape: one WARP per candidate (e.g. n=100, d=10: 10 KB).
//
// Replaces the same reference code as the other search kernels: generate_lhd
// (src/lhd.rs:97-132) -> criterion (src/maxpro_utils.rs:28-83 / src/maximin_utils.rs:54-67)
// -> fold (src/lib.rs:76-98).  It fills the gap between the thread-group kernels
// (search_small.cu: designs up to 3.5 KB) and the CTA-per-candidate kernel (search_cta.cu):
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
// common.cuh (candidates are in the unit cube by construction).
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cstdlib>

namespace mp {

constexpr int kWarpCtaThreads = 256;  // 8 candidates per CTA
constexpr int kWarpSB = 8;            // shifts per work item

struct WarpParams {
    int n, d, ns;  // ns = odd row stride of the [k][i] layout (the column walkers hit different banks)
    unsigned int cand_bytes;  // shared memory of one candidate slot
    unsigned long long seed, lo, hi;
    double c1, c0, n_pairs, inv_d;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

template <bool MAXPRO>
__global__ void __launch_bounds__(kWarpCtaThreads, 2) search_warp_kernel(WarpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BestRec warp_best[kWarpCtaThreads / 32];
    const int n = p.n, d = p.d, ns = p.ns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = reinterpret_cast<double*>(smem_raw + (size_t)warp * p.cand_bytes);  // [d][ns]
    unsigned short* jst = reinterpret_cast<unsigned short*>(xs + (size_t)d * ns);    // [d][ns]
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
            jst[k * ns + i] = (unsigned short)mulhi_bound(w.y, (uint32_t)i + 1u);
            if (i + 1 < n) {
                xs[k * ns + i + 1] = stratum_value((uint32_t)(i + 1), w.z, p.c1, p.c0);
                jst[k * ns + i + 1] = (unsigned short)mulhi_bound(w.w, (uint32_t)i + 2u);
            }
        }
        __syncwarp();
        // ---- generation, phase 2: one lane per column
        for (int k = lane; k < d; k += 32) {
            double* a = xs + k * ns;
            const unsigned short* js = jst + k * ns;
            for (int i = 0; i < n; ++i) {
                const int j = js[i];
                const double v = a[i];
                const double t = a[j];
                a[i] = t;
                a[j] = v;
            }
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
            int jj[kWarpSB];
#pragma unroll
            for (int c = 0; c < kWarpSB; ++c) {
                int j = i + s0 + c;
                if (j >= n) j -= n;
                jj[c] = (s0 + c <= my_smax) ? j : i;  // masked slots pair the row with itself
            }
            double pr[kWarpSB];
            {
                const double xi = xs[i];
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) pr[c] = xi - xs[jj[c]];
                    else { const double df = __dsub_rn(xi, xs[jj[c]]); pr[c] = __dmul_rn(df, df); }
                }
            }
            for (int k = 1; k < d; ++k) {
                const double* xk = xs + k * ns;
                const double xi = xk[i];
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) pr[c] *= xi - xk[jj[c]];
                    else { const double df = __dsub_rn(xi, xk[jj[c]]); pr[c] = __dadd_rn(pr[c], __dmul_rn(df, df)); }
                }
            }
            if (MAXPRO) {
                if (pushed == kMaxChainTerms) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) acc += ch[c].flush();
                    pushed = 0;
                }
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c)
                    if (s0 + c <= my_smax) ch[c].push(fma(pr[c], pr[c], kEps));
                ++pushed;
            } else {
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c)
                    if (s0 + c <= my_smax) mn = fmin(mn, pr[c]);
            }
            // next item of this lane: 32 rows further, wrapping into the next shift block
            i += 32;
            while (i >= n) { i -= n; ++b; }
        }
        if (MAXPRO) {
#pragma unroll
            for (int c = 0; c < kWarpSB; ++c) acc += ch[c].flush();
        }
        double s = MAXPRO ? warp_reduce_sum(acc) : warp_reduce_min(mn);
        if (!MAXPRO) s = sqrt(s);
        if (MAXPRO ? (s <= best) : (s >= best)) { best = s; best_idx = idx; }  // later index on ties
    }
    if (MAXPRO && best_idx != kNoIndex) best = pow(best / p.n_pairs, p.inv_d);  // maxpro_utils.rs:81-82
    if (lane != 0) { best = MAXPRO ? CUDART_INF : -CUDART_INF; best_idx = kNoIndex; }
    block_publish_best<MAXPRO>(best, best_idx, warp_best, p.block_best, p.done_counter, p.result);
}

static unsigned int warp_cand_bytes(unsigned long long n, unsigned long long d) {
    const unsigned long long ns = n | 1ull;
    return (unsigned int)(((d * ns * 10 + 15) / 16) * 16);
}

// Medium designs: too big for the thread-group kernels, small enough that 16 candidates fit the
// shared memory of one SM (two CTAs of 8 warps).
bool search_warp_supported(unsigned long long n, unsigned long long d) {
    if (const char* e = getenv("MAXPRO_NO_WARP")) if (atoi(e) != 0) return false;
    return n >= 2 && n <= 65535 && d >= 1 && (size_t)warp_cand_bytes(n, d) * 16 + 2048 <= kSmemBudget;
}

int search_warp_grid(int sms) { return 2 * sms; }

cudaError_t launch_search_warp(const SearchLaunch& L) {
    WarpParams p;
    p.n = (int)L.n; p.d = (int)L.d; p.ns = (int)(L.n | 1ull);
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

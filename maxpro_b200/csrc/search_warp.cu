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
// of the lower half for even n), as work items of (two neighbouring rows, 8 consecutive shifts):
// 16 pairs from six 16-byte loads per dimension -- the two rows in one, their ten partner rows in
// five (0.4 loads per pair; one row per item needed 1.1).  The items are dealt to the lanes in
// row-major order (consecutive lanes <-> consecutive row pairs: contiguous, conflict-free), so
// every lane gets the same number of items whatever n is.  The 16 running products stay in
// registers across the dimensions; MaxPro terms go through the reciprocal chains of common.cuh
// (candidates are in the unit cube by construction).  Each column is followed by a copy of its
// first 8*ceil(smax/8)+2 rows, so row (i+s) mod n is simply row i+s: the partner loads of an item
// are one address plus immediates, with no wrap arithmetic and no index array.
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cstdlib>

namespace mp {

constexpr int kWarpCtaThreads = 128;  // 4 candidates per CTA, up to 4 CTAs per SM
constexpr int kWarpSB = 8;            // shifts per work item

struct WarpParams {
    int gen;       // design stream version (philox.cuh)
    int n, d, ns;  // ns = row stride of the [k][i] layout: n + copied rows, even (16-byte loads), 2 or 10 mod 16 (the column walkers hit different banks)
    unsigned int cand_bytes;  // shared memory of one candidate slot
    unsigned long long seed, lo, hi;
    // list mode (stage 2 of the two-stage search, search_screen.cu): score list[0..*list_count_dev) instead of lo..hi; a count
    // above list_cap means stage 1 overflowed its list: the whole range is scored instead
    const unsigned long long* list;
    const unsigned int* list_count_dev;
    unsigned int list_cap;
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
    const int npairs = (n + 1) / 2;            // row pairs (the last one has a single row when n is odd)
    const int items = npairs * nblk;
    const int wrap = kWarpSB * nblk + 2;       // copied rows behind each column
    const int ns2 = ns / 2;                    // row stride in 16-byte units

    LaneBest<MAXPRO> lb;  // fold of src/lib.rs:85-97 on the criterion (reduce.cuh); every lane of the warp holds the same state
    lb.init(p.inv_d);

    const unsigned long long wstep = (unsigned long long)gridDim.x * (kWarpCtaThreads / 32);
    const unsigned long long* list = p.list;
    unsigned long long list_count = 0;
    if (list) {
        list_count = *p.list_count_dev;
        if (list_count > p.list_cap) list = nullptr;
    }
    const unsigned long long total = list ? list_count : p.hi - p.lo;
    for (unsigned long long off = (unsigned long long)blockIdx.x * (kWarpCtaThreads / 32) + warp; off < total; off += wstep) {
        const unsigned long long idx = list ? list[off] : p.lo + off;
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        __syncwarp();  // previous candidate fully consumed
        // ---- generation, phase 1: block q <-> (column k, element pair b), walked without a division
        {
            int k = 0, b = lane;
            while (b >= nb) { b -= nb; ++k; }
            LhdKeyCache kc{0xFFFFFFFFu, 0u, 0u};
            while (k < d) {
                const LhdBlock w = lhd_draw_block(p.gen, (uint32_t)b, (uint32_t)k, s_lo, s_hi, kc);
                const int i = 2 * b;
                double* col = xs + k * ns + i;
                unsigned short* jc = jst + k * n + i;
                col[0] = lhd_value(p.gen, (uint32_t)i, w.jit0, p.c1, p.c0);
                jc[0] = (unsigned short)w.j0;
                if (i + 1 < n) {
                    col[1] = lhd_value(p.gen, (uint32_t)(i + 1), w.jit1, p.c1, p.c0);
                    jc[1] = (unsigned short)w.j1;
                }
                b += 32;
                while (b >= nb) { b -= nb; ++k; }
            }
        }
        __syncwarp();
        // ---- generation, phase 2: one lane per column
        for (int k = lane; k < d; k += 32) fy_walk_column(xs + k * ns, jst + k * n, n);
        __syncwarp();
        // rows n .. n + wrap - 1 repeat rows 0 .. (wrapping again if the design is tiny)
        for (int k = 0; k < d; ++k) {
            double* col = xs + k * ns;
            for (int r = lane; r < wrap; r += 32) {
                int src = r;
                while (src >= n) src -= n;
                col[n + r] = col[src];
            }
        }
        __syncwarp();

        // ---- scoring: item q <-> (row pair q mod npairs, shift block q / npairs)
        double acc = 0.0, mn = CUDART_INF;
        RecipChain ch[kWarpSB];
#pragma unroll
        for (int c = 0; c < kWarpSB; ++c) ch[c].reset();
        int pushed = 0;
        int m = lane, b = 0;
        while (m >= npairs) { m -= npairs; ++b; }  // n may be below 64
        for (int q = lane; q < items; q += 32) {
            const int i0 = 2 * m, i1 = i0 + 1;
            const int s0 = 1 + b * kWarpSB;
            // rows i0, i0+1 (one 16-byte load) against rows i0 + s0 - 1 .. i0 + s0 + 8 (five 16-byte loads):
            // row i0 meets partners [1, 9), row i0+1 partners [2, 10) -- the same eight shifts s0 .. s0+7
            const double2* pa = reinterpret_cast<const double2*>(xs + i0);
            const double2* pp = reinterpret_cast<const double2*>(xs + i0 + s0 - 1);
            double pr[2][kWarpSB];
            {
                const double2 a = pa[0];
                double pt[kWarpSB + 2];
#pragma unroll
                for (int c = 0; c < kWarpSB / 2 + 1; ++c) { const double2 v = pp[c]; pt[2 * c] = v.x; pt[2 * c + 1] = v.y; }
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) { pr[0][c] = a.x - pt[c + 1]; pr[1][c] = a.y - pt[c + 2]; }
                    else {
                        const double d0 = __dsub_rn(a.x, pt[c + 1]), d1 = __dsub_rn(a.y, pt[c + 2]);
                        pr[0][c] = __dmul_rn(d0, d0); pr[1][c] = __dmul_rn(d1, d1);
                    }
                }
            }
            for (int k = 1; k < d; ++k) {
                pa += ns2; pp += ns2;
                const double2 a = pa[0];
                double pt[kWarpSB + 2];
#pragma unroll
                for (int c = 0; c < kWarpSB / 2 + 1; ++c) { const double2 v = pp[c]; pt[2 * c] = v.x; pt[2 * c + 1] = v.y; }
#pragma unroll
                for (int c = 0; c < kWarpSB; ++c) {
                    if (MAXPRO) { pr[0][c] *= a.x - pt[c + 1]; pr[1][c] *= a.y - pt[c + 2]; }
                    else {
                        const double d0 = __dsub_rn(a.x, pt[c + 1]), d1 = __dsub_rn(a.y, pt[c + 2]);
                        pr[0][c] = __dadd_rn(pr[0][c], __dmul_rn(d0, d0));
                        pr[1][c] = __dadd_rn(pr[1][c], __dmul_rn(d1, d1));
                    }
                }
            }
            // shift s0 + c of row i is a pair iff s0 + c <= half, or it is the diameter of a lower-half row
            const int smax0 = half + ((even && i0 < n / 2) ? 1 : 0);
            const int smax1 = i1 < n ? half + ((even && i1 < n / 2) ? 1 : 0) : 0;  // odd n: the last item has one row
            const bool whole = s0 + kWarpSB - 1 <= half && i1 < n;
            if (MAXPRO) {
                if (pushed + 2 > kMaxChainTerms) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) acc = ch[c].flush_into(acc);
                    pushed = 0;
                }
                if (whole) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) {
                        ch[c].push(fma(pr[0][c], pr[0][c], kEps));
                        ch[c].push(fma(pr[1][c], pr[1][c], kEps));
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) {
                        if (s0 + c <= smax0) ch[c].push(fma(pr[0][c], pr[0][c], kEps));
                        if (s0 + c <= smax1) ch[c].push(fma(pr[1][c], pr[1][c], kEps));
                    }
                }
                pushed += 2;
            } else {
                if (whole) {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) mn = min_sqdist(mn, min_sqdist(pr[0][c], pr[1][c]));
                } else {
#pragma unroll
                    for (int c = 0; c < kWarpSB; ++c) {
                        if (s0 + c <= smax0) mn = min_sqdist(mn, pr[0][c]);
                        if (s0 + c <= smax1) mn = min_sqdist(mn, pr[1][c]);
                    }
                }
            }
            // next item of this lane: 32 row pairs further, wrapping into the next shift block
            m += 32;
            while (m >= npairs) { m -= npairs; ++b; }
        }
        if (MAXPRO) {
#pragma unroll
            for (int c = 0; c < kWarpSB; ++c) acc = ch[c].flush_into(acc);
        }
        double s = MAXPRO ? warp_reduce_sum(acc) : warp_reduce_min(mn);
        if (!MAXPRO) s = sqrt(s);
        lb.offer(s, idx, p.n_pairs, p.inv_d);
    }
    if (lane != 0) lb.init(p.inv_d);  // one record per warp
    block_publish_best<MAXPRO>(lb.best, lb.idx, warp_best, p.block_best, p.done_counter, p.result);
}

static unsigned long long warp_row_stride(unsigned long long n) {
    const unsigned long long smax = (n - 1) / 2 + ((n & 1) == 0 ? 1 : 0);
    const unsigned long long nblk = (smax + kWarpSB - 1) / kWarpSB;
    unsigned long long ns = n + kWarpSB * nblk + 2;  // n rows + the copied ones
    ns += ns & 1;                                     // even: every column starts on a 16-byte boundary
    while (ns % 16 != 2 && ns % 16 != 10) ns += 2;    // the lanes that walk one column each fall into different bank groups
    return ns;
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
    if (const char* e = cfg("MAXPRO_NO_WARP")) if (atoi(e) != 0) return false;
    int need = 3;
    if (const char* e = cfg("MAXPRO_WARP_MIN_CTAS")) { int v = atoi(e); if (v >= 1 && v <= 4) need = v; }
    return n >= 2 && n <= 65535 && d >= 1 && d <= 65535 && warp_ctas_per_sm(n, d) >= need;
}

int search_warp_grid(unsigned long long n, unsigned long long d, int sms) { return warp_ctas_per_sm(n, d) * sms; }

cudaError_t launch_search_warp(const SearchLaunch& L) {
    WarpParams p;
    p.n = (int)L.n; p.d = (int)L.d; p.ns = (int)warp_row_stride(L.n);
    p.cand_bytes = warp_cand_bytes(L.n, L.d);
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.list = L.list; p.list_count_dev = L.list_count_dev; p.list_cap = L.list_cap;
    p.gen = stream_version();
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = lhd_c0(p.gen, p.c1);
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

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

namespace mp {

constexpr int kCtaThreads = 128;
constexpr int kTileRows = 64;  // 2 rows per lane
constexpr int kJB = 8;         // columns per block

struct CtaParams {
    int n, d, ns;  // ns = row stride of the [k][i] layout (cta_row_stride)
    unsigned long long seed, lo, hi;
    double c1, c0, n_pairs, inv_d;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

// Eight consecutive partner coordinates as four 16-byte warp-wide broadcasts.  The row stride is
// even and a block starts at a multiple of 8, so the address is 16-byte aligned; a block that
// runs past row n-1 reads the stride padding (or the next dimension) and is masked by the caller.
__device__ __forceinline__ void load_block(double (&bv)[kJB], const double* src) {
#pragma unroll
    for (int c = 0; c < kJB; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + c);
        bv[c] = v.x; bv[c + 1] = v.y;
    }
}

// Row stride of the [k][i] layout: n rounded up to a multiple of 8 (whole partner blocks), plus 2.
// Even, so partner blocks are 16-byte aligned; 2 or 10 modulo 16, so the lanes that walk one
// column each in generation phase 2 (stride ns doubles) fall into 8 different bank groups.
__host__ __device__ __forceinline__ unsigned long long cta_row_stride(unsigned long long n) { return (n + 7) / 8 * 8 + 2; }

template <bool MAXPRO>
__global__ void __launch_bounds__(kCtaThreads, 4) search_cta_kernel(CtaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, d = p.d, ns = p.ns;
    double* xs = reinterpret_cast<double*>(smem_raw);                         // [d][ns]
    unsigned short* jst = reinterpret_cast<unsigned short*>(xs + (size_t)d * ns);  // [d][ns]
    double* red = reinterpret_cast<double*>(smem_raw + (((size_t)d * ns * 10 + 15) / 16) * 16);  // [4] warp partials
    __shared__ BestRec warp_best[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = (n + 1) / 2;
    const int tiles = (n + kTileRows - 1) / kTileRows;

    double best = MAXPRO ? CUDART_INF : -CUDART_INF;  // kept by thread 0
    unsigned long long best_idx = kNoIndex;

    for (unsigned long long idx = p.lo + blockIdx.x; idx < p.hi; idx += gridDim.x) {
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        __syncthreads();  // previous candidate fully consumed
        // ---- generation, phase 1
        for (int q = tid; q < d * nb; q += kCtaThreads) {
            const int k = q / nb, b = q - k * nb;
            Philox4 w = philox4x32_10((uint32_t)b, (uint32_t)k, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
            const int i = 2 * b;
            xs[k * ns + i] = fma((double)(((unsigned long long)(uint32_t)i << 32) | w.x), p.c1, p.c0);
            jst[k * ns + i] = (unsigned short)mulhi_bound(w.y, (uint32_t)i + 1u);
            if (i + 1 < n) {
                xs[k * ns + i + 1] = fma((double)(((unsigned long long)(uint32_t)(i + 1) << 32) | w.z), p.c1, p.c0);
                jst[k * ns + i + 1] = (unsigned short)mulhi_bound(w.w, (uint32_t)i + 2u);
            }
        }
        __syncthreads();
        // ---- generation, phase 2: one lane per column
        for (int k = tid; k < d; k += kCtaThreads) {
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
        __syncthreads();

        // ---- scoring
        double acc = 0.0, mn = CUDART_INF;
        RecipChain ch[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) ch[c].reset();
        int pushed = 0;
        int item = 0;
        for (int t = 0; t < tiles; ++t) {
            const int row0 = t * kTileRows;
            const int i0 = row0 + lane, i1 = i0 + 32;
            const int i0c = min(i0, n - 1), i1c = min(i1, n - 1);
            const int nblk = (n - row0 + kJB - 1) / kJB;  // column blocks from the tile's first row
            for (int jb = 0; jb < nblk; ++jb, ++item) {
                if ((item & 3) != warp) continue;
                const int j0 = row0 + jb * kJB;
                double prod[2][kJB];
                // dimension 0, then the rest
                {
                    const double a0 = xs[i0c], a1 = xs[i1c];
                    double bv[kJB];
                    load_block(bv, xs + j0);
#pragma unroll
                    for (int c = 0; c < kJB; ++c) {
                        const double b = bv[c];
                        if (MAXPRO) { prod[0][c] = a0 - b; prod[1][c] = a1 - b; }
                        else {
                            double d0 = __dsub_rn(a0, b), d1 = __dsub_rn(a1, b);
                            prod[0][c] = __dmul_rn(d0, d0); prod[1][c] = __dmul_rn(d1, d1);
                        }
                    }
                }
                for (int k = 1; k < d; ++k) {
                    const double* xk = xs + k * ns;
                    const double a0 = xk[i0c], a1 = xk[i1c];
                    double bv[kJB];
                    load_block(bv, xk + j0);
#pragma unroll
                    for (int c = 0; c < kJB; ++c) {
                        const double b = bv[c];
                        if (MAXPRO) { prod[0][c] *= (a0 - b); prod[1][c] *= (a1 - b); }
                        else {
                            double d0 = __dsub_rn(a0, b), d1 = __dsub_rn(a1, b);
                            prod[0][c] = __dadd_rn(prod[0][c], __dmul_rn(d0, d0));
                            prod[1][c] = __dadd_rn(prod[1][c], __dmul_rn(d1, d1));
                        }
                    }
                }
                // full blocks: every row of the tile precedes every column and all are < n
                const bool full = (j0 >= row0 + kTileRows) && (j0 + kJB <= n) && (row0 + kTileRows <= n);
                if (MAXPRO) {
                    if (pushed + 2 > kMaxChainTerms) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc += ch[c].flush();
                        pushed = 0;
                    }
                    if (full) {
#pragma unroll
                        for (int c = 0; c < kJB; ++c) {
                            ch[c].push(fma(prod[0][c], prod[0][c], kEps));
                            ch[c].push(fma(prod[1][c], prod[1][c], kEps));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < kJB; ++c) {
                            const int j = j0 + c;
                            if (j < n && i0 < j) ch[c].push(fma(prod[0][c], prod[0][c], kEps));
                            if (j < n && i1 < j) ch[c].push(fma(prod[1][c], prod[1][c], kEps));
                        }
                    }
                    pushed += 2;
                } else {
#pragma unroll
                    for (int c = 0; c < kJB; ++c) {
                        const int j = j0 + c;
                        if (full || (j < n && i0 < j)) mn = fmin(mn, prod[0][c]);
                        if (full || (j < n && i1 < j)) mn = fmin(mn, prod[1][c]);
                    }
                }
            }
        }
        if (MAXPRO) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc += ch[c].flush();
        }
        double part = MAXPRO ? warp_reduce_sum(acc) : warp_reduce_min(mn);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            double s = red[0];
#pragma unroll
            for (int w = 1; w < 4; ++w) s = MAXPRO ? s + red[w] : fmin(s, red[w]);
            if (!MAXPRO) s = sqrt(s);
            if (MAXPRO ? (s <= best) : (s >= best)) { best = s; best_idx = idx; }  // later index on ties
        }
    }
    if (tid == 0 && MAXPRO && best_idx != kNoIndex) best = pow(best / p.n_pairs, p.inv_d);  // maxpro_utils.rs:81-82
    if (tid != 0) { best = MAXPRO ? CUDART_INF : -CUDART_INF; best_idx = kNoIndex; }
    block_publish_best<MAXPRO>(best, best_idx, warp_best, p.block_best, p.done_counter, p.result);
}

static size_t cta_smem_bytes(unsigned long long n, unsigned long long d) {
    unsigned long long ns = cta_row_stride(n);
    return (size_t)(((d * ns * 10 + 15) / 16) * 16 + 4 * sizeof(double));
}

bool search_cta_supported(unsigned long long n, unsigned long long d) {
    return n >= 2 && n <= 65535 && d >= 1 && cta_smem_bytes(n, d) + 256 <= kSmemBudget;  // 256: static shared memory
}

int search_cta_grid(unsigned long long n, unsigned long long d, int sms) {
    size_t smem = cta_smem_bytes(n, d) + 1024;
    int per_sm = (int)(kSmemBudget / smem);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

cudaError_t launch_search_cta(const SearchLaunch& L) {
    CtaParams p;
    p.n = (int)L.n; p.d = (int)L.d; p.ns = (int)cta_row_stride(L.n);
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = 0.5 * p.c1;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.block_best = L.block_best; p.done_counter = L.done_counter; p.result = L.result;
    size_t smem = cta_smem_bytes(L.n, L.d);
    cudaError_t e;
    if (L.metric == 0) {
        e = cudaFuncSetAttribute(search_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_cta_kernel<true><<<L.grid, kCtaThreads, smem, L.stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(search_cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        search_cta_kernel<false><<<L.grid, kCtaThreads, smem, L.stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace mp

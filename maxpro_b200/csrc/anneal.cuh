// Device helpers shared by the annealing kernels (anneal.cu: every mode; anneal_pipe.cu: the
// software-pipelined swap/MaxPro chain).  See anneal.cu for the reference mapping.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "kernels.h"
#include "tile_score.cuh"

namespace mp {

constexpr int kDrawBatch = 128;  // swap iterations whose draws are generated together
constexpr int kShiftBlock = 8;   // cyclic shifts per full-score work item

struct AnnealParams {
    const double* design;  // [n][d] row-major start design (shared by all chains)
    unsigned long long n, d, iters, seed, n_chains;
    double t0, cooling, step;
    int minimize, swap;
    unsigned long long resync;
    double n_pairs, inv_d;
    double* best_design;        // [grid][2][n*d] row-major: two buffers per CTA, see CtaBest
    double* best_metric;        // [n_chains]
    unsigned int* improved;     // [n_chains]
    unsigned long long* cta_chain;  // [grid] the chain whose global best this CTA keeps (kNoIndex: none improved)
    int* cta_buf;                   // [grid] which of the CTA's two buffers holds it
    const double* raw0;         // S (or min squared distance) of the start design: every chain starts there,
                                // so it is computed once by anneal_raw0_kernel instead of once per chain
    double* cur_global;         // anneal.cu, jitter moves on designs that fit shared memory only once: [grid][d][ns] scratch
    const double* rowmin0;      // anneal_maximin.cu: nearest-neighbour squared distance of every row of the start design
    const unsigned short* rowarg0;  //               and the neighbour that attains it
    const double* qparams;      // anneal_maximin.cu, screened kernel: colmin[d], quantisation step s, 1/s, usable flag
};

// Per-CTA fold of the chains' global bests.  The chains of a launch are grid-strided over the CTAs, and the caller
// wants ONE design back -- the best global best over all chains, the lowest chain index on equal metrics.  So a CTA
// owns two design buffers in HBM: the chain in progress writes its global best (anneal.rs:131-136) into one, the
// other keeps the best finished chain of this CTA.  A finished chain that improved on its start and is STRICTLY
// better than the kept one takes its place by flipping the two roles (the chains of a CTA run in increasing index
// order, so the earlier chain survives a tie).  Scratch is 2 * grid * n * d doubles instead of n_chains * n * d:
// 47 MB instead of 10.5 GB for 65,536 chains of a 1000 x 20 design.
struct CtaBest {
    double metric;
    unsigned long long chain;
    int scratch, has;
};
__device__ __forceinline__ void cta_best_reset(CtaBest* st, const AnnealParams& p) {  // one thread, before the first chain
    st->metric = 0.0; st->chain = kNoIndex; st->scratch = 0; st->has = 0;
    p.cta_chain[blockIdx.x] = kNoIndex;
    p.cta_buf[blockIdx.x] = 0;
}
// where the chain in progress writes its global best; read after the barrier at the top of a chain
__device__ __forceinline__ double* cta_best_slot(const AnnealParams& p, const CtaBest* st, int nd) {
    return p.best_design + ((size_t)blockIdx.x * 2 + (size_t)st->scratch) * (size_t)nd;
}
// one thread, when a chain ends
__device__ __forceinline__ void cta_best_commit(CtaBest* st, const AnnealParams& p, unsigned long long chain, double global_best, bool materialized) {
    p.best_metric[chain] = global_best;
    p.improved[chain] = materialized ? 1u : 0u;
    if (materialized && (!st->has || (p.minimize ? global_best < st->metric : global_best > st->metric))) {
        st->has = 1; st->metric = global_best; st->chain = chain;
        p.cta_chain[blockIdx.x] = chain;
        p.cta_buf[blockIdx.x] = st->scratch;
        st->scratch ^= 1;
    }
}

// Host: grid of an anneal launch = resident CTAs of `kernel` (after its shared-memory attribute is set), one per chain at most.
template <class K>
inline cudaError_t anneal_resident_grid(K kernel, int threads, size_t smem, int sms, unsigned long long n_chains, unsigned long long cap, unsigned int* grid) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    unsigned long long g = (unsigned long long)(per_sm < 1 ? 1 : per_sm) * (unsigned long long)(sms < 1 ? 1 : sms);
    if (g > n_chains) g = n_chains;
    if (g > cap) g = cap;
    *grid = (unsigned int)g;
    return cudaSuccess;
}

// Deterministic block reduction in two halves.  Every thread: lane tree, lane 0 publishes the
// warp's partial.  After a barrier, warp 0 alone folds the partials (a second lane tree) --
// the chain's scalar state and the Metropolis arithmetic (pow, exp, two divides) live in warp
// 0 only, so the other warps spend no issue slots on it.
template <bool IS_SUM>
__device__ __forceinline__ void publish_partial(double v, double* part) {
    v = IS_SUM ? warp_reduce_sum(v) : warp_reduce_min(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
}
template <bool IS_SUM>
__device__ __forceinline__ double warp_total(const double* part, int nw) {  // nw <= 32 partials
    const int lane = threadIdx.x & 31;
    double v = lane < nw ? part[lane] : (IS_SUM ? 0.0 : CUDART_INF);
    return IS_SUM ? warp_reduce_sum(v) : warp_reduce_min(v);
}

// Full criterion of the design at xs ([k][i] layout, row stride ns).  MAXPRO: S
// (maxpro_utils.rs:38-58); else min squared distance, bit-exact per pair
// (maximin_utils.rs:39-44).  Pairs are enumerated cyclically -- row i meets rows
// i+1 .. i+half (mod n), and for even n the diameters (i, i+n/2) belong to the lower
// half -- in items of kShiftBlock consecutive shifts whose partner rows are consecutive
// addresses across the lanes.  Outside the unit cube (in_cube false) every term takes an
// IEEE divide instead of the reciprocal chains.
template <bool MAXPRO, int SB = kShiftBlock>
__device__ __forceinline__ double full_score_partial(const double* xs, int n, int d, int ns, bool in_cube) {
    double acc = 0.0, mn = CUDART_INF;
    RecipChain ch[SB];
#pragma unroll
    for (int c = 0; c < SB; ++c) ch[c].reset();
    int pushed = 0;
    const int half = (n - 1) / 2;
    const bool even = (n & 1) == 0;
    const int smax = half + (even ? 1 : 0);
    // Work items (row i, SB consecutive shifts from s0), dealt row-major to the threads: consecutive threads take
    // consecutive rows of the same shift block, and a block of more than n threads works on several shift blocks
    // at once (a 50-row design has 200 items: one per thread instead of four per row-thread in sequence).
    const int nitems = (smax + SB - 1) / SB;
    for (int w = threadIdx.x; w < n * nitems; w += blockDim.x) {
        const int it = w / n, i = w - it * n;
        const int my_smax = half + ((even && i < n / 2) ? 1 : 0);
        {
            const int s0 = 1 + it * SB;
            int jj[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                int j = i + s0 + c;
                if (j >= n) j -= n;
                jj[c] = (s0 + c <= my_smax) ? j : i;  // masked slots pair the row with itself
            }
            double q[SB];
            {
                const double xi = xs[i];
#pragma unroll
                for (int c = 0; c < SB; ++c) {
                    if (MAXPRO) q[c] = xi - xs[jj[c]];
                    else { const double df = __dsub_rn(xi, xs[jj[c]]); q[c] = __dmul_rn(df, df); }
                }
            }
            for (int k = 1; k < d; ++k) {
                const double* xk = xs + k * ns;
                const double xi = xk[i];
#pragma unroll
                for (int c = 0; c < SB; ++c) {
                    if (MAXPRO) q[c] *= xi - xk[jj[c]];
                    else { const double df = __dsub_rn(xi, xk[jj[c]]); q[c] = __dadd_rn(q[c], __dmul_rn(df, df)); }
                }
            }
            if (MAXPRO) {
                if (in_cube) {
                    if (pushed == kMaxChainTerms) {
#pragma unroll
                        for (int c = 0; c < SB; ++c) acc = ch[c].flush_into(acc);
                        pushed = 0;
                    }
#pragma unroll
                    for (int c = 0; c < SB; ++c)
                        if (s0 + c <= my_smax) ch[c].push(fma(q[c], q[c], kEps));
                    ++pushed;
                } else {
#pragma unroll
                    for (int c = 0; c < SB; ++c)
                        if (s0 + c <= my_smax) acc += 1.0 / fma(q[c], q[c], kEps);
                }
            } else {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (s0 + c <= my_smax) mn = min_sqdist(mn, q[c]);
            }
        }
    }
    if (MAXPRO) {
#pragma unroll
        for (int c = 0; c < SB; ++c) acc = ch[c].flush_into(acc);
    }
    return MAXPRO ? acc : mn;  // this thread's share; the caller reduces
}

// The same criterion over work items of 64 rows x 8 columns (tile_score.cuh), dealt round-robin to the
// warps: 8 shared-memory wavefronts per dimension for 16 pairs per lane, where the cyclic items above
// take 18 for 8 -- and the full re-score is bound by exactly that (16 warps x 18 wavefronts against
// 132 cycles of FP64 issue per dimension step).  Worth it from n ~ 100 up; below that the 64-row
// tiles are mostly padding.  xs must be 16-byte aligned with an even row stride; the last column
// block may read up to 6 doubles past row n - 1 of the last dimension (masked), so the buffer needs
// that much slack behind it.
template <bool MAXPRO>
__device__ __forceinline__ double full_score_tiles(const double* xs, int n, int d, int ns, bool in_cube) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int tiles = (n + kTileRows - 1) / kTileRows;
    const int chunk = cta_dim_chunk(d);
    CtaScore st;
    st.reset();
    int t = 0, jb = warp;
    int nblk = (n + kJB - 1) / kJB;  // column blocks from the tile's first row
    while (true) {
        while (t < tiles && jb >= nblk) { jb -= nblk; ++t; nblk = (n - t * kTileRows + kJB - 1) / kJB; }
        if (t >= tiles) break;
        if (!MAXPRO || in_cube) cta_score_item_any<MAXPRO, false>(chunk, xs, n, d, ns, t * kTileRows, t * kTileRows + jb * kJB, lane, st);
        else cta_score_item_any<MAXPRO, true>(chunk, xs, n, d, ns, t * kTileRows, t * kTileRows + jb * kJB, lane, st);
        jb += nw;
    }
    if (MAXPRO) st.flush();
    return MAXPRO ? st.acc : st.mn;  // this thread's share; the caller reduces
}

// Designs from this size up take full_score_tiles.  Measured with 1,184 jitter chains (algorithmic TFLOP/s,
// cyclic items -> tiles): (100,5) 5.4 -> 6.9, (128,8) 8.7 -> 10.3, (200,10) 7.6 -> 11.5, (250,4) 8.2 -> 11.5,
// (1000,20) 10.6 -> 18.9; (64,6) 5.7 -> 5.4.  A single (100,5) chain: 5.5 -> 4.9 us per step.
#ifndef MAXPRO_TILE_SCORE_MIN_N
#define MAXPRO_TILE_SCORE_MIN_N 96
#endif
constexpr int kTileScoreMinN = MAXPRO_TILE_SCORE_MIN_N;

template <bool MAXPRO>
__device__ __forceinline__ double metric_from_raw(double raw, int n, double n_pairs, double inv_d) {
    if (MAXPRO) {  // maxpro_utils.rs:75-82
        if (n < 2) return 0.0;
        const double m = raw / n_pairs;
        // pow costs ~800 cycles of dependent latency on the one warp that decides (tools/micro/lat.cu).  x^1 is x;
        // x^0.5 is sqrt(x), the correctly rounded value of what powf(0.5) approximates.  d = 2 is the CLI's own case.
        if (inv_d == 1.0) return m;
        if (inv_d == 0.5) return sqrt(m);
        return pow(m, inv_d);
    }
    return sqrt(raw);                                              // maximin_utils.rs:44
}

// Change of S when x[r1][c] and x[r2][c] are exchanged: the partial sum of this thread over
// its partner rows.  Per partner row j (not r1, r2):
//   A = prod_{k != c} (x[r1][k] - x[j][k]),  B likewise for r2;  u = a_c - x[j][c], v = b_c - x[j][c]
//   pair (r1, j): p = (A u)^2 + eps -> p' = (A v)^2 + eps;   pair (r2, j): p = (B v)^2 + eps -> p' = (B u)^2 + eps
//   1/p' - 1/p = (p - p') / (p p')
// e1, e2: two more partner rows to leave out (-1 = none); `workers` threads share the rows.
__device__ __forceinline__ double swap_delta_partial(const double* cur, const double2* rows, int n, int d, int ns,
                                                     int r1, int r2, int c, bool in_cube, int e1, int e2, int workers) {
    // rows[k] = {x[r1][k], x[r2][k]}: staged once per iteration, read as 16-byte broadcasts
    double delta = 0.0;
    const double a_c = rows[c].x, b_c = rows[c].y;
    for (int j0 = 2 * threadIdx.x; j0 < n; j0 += 2 * workers) {
        const double* qx = cur + j0;
        const double2 xc = *reinterpret_cast<const double2*>(qx + c * ns);
        double A0 = 1.0, B0 = 1.0, A1 = 1.0, B1 = 1.0;
        const double2* rw = rows;
#pragma unroll 4
        for (int k = 0; k < c; ++k, qx += ns, ++rw) {
            const double2 x = *reinterpret_cast<const double2*>(qx);
            const double2 ab = *rw;
            A0 *= ab.x - x.x; B0 *= ab.y - x.x; A1 *= ab.x - x.y; B1 *= ab.y - x.y;
        }
        qx += ns; ++rw;
#pragma unroll 4
        for (int k = c + 1; k < d; ++k, qx += ns, ++rw) {
            const double2 x = *reinterpret_cast<const double2*>(qx);
            const double2 ab = *rw;
            A0 *= ab.x - x.x; B0 *= ab.y - x.x; A1 *= ab.x - x.y; B1 *= ab.y - x.y;
        }
        double nm[2], dn[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = j0 + h;
            const double A = h ? A1 : A0, B = h ? B1 : B0, xjc = h ? xc.y : xc.x;
            const double uu = a_c - xjc, vv = b_c - xjc;
            const double au = A * uu, av = A * vv, bu = B * uu, bv = B * vv;
            const double p1 = fma(au, au, kEps), p1n = fma(av, av, kEps);
            const double p2 = fma(bv, bv, kEps), p2n = fma(bu, bu, kEps);
            const bool live = j != r1 && j != r2 && j != e1 && j != e2 && j < n;
            if (in_cube) {
                // both changes over one denominator: (p1-p1')/(p1 p1') + (p2-p2')/(p2 p2')
                const double d1 = p1 * p1n, d2 = p2 * p2n;
                nm[h] = live ? fma(p1 - p1n, d2, (p2 - p2n) * d1) : 0.0;
                dn[h] = live ? d1 * d2 : 1.0;  // >= 1e-48 in the unit cube
            } else if (live) {
                delta += (p1 - p1n) / (p1 * p1n) + (p2 - p2n) / (p2 * p2n);
            }
        }
        if (in_cube) delta += fma(nm[0], dn[1], nm[1] * dn[0]) / (dn[0] * dn[1]);  // denominator >= 1e-96
    }
    return delta;
}

// What warp 0 tells the other warps about an iteration.
struct Decision {
    int accept;
    int best_mode;  // 0 none, 1 replay the move log into the HBM best, 2 rewrite the HBM best in full
    int log_n;      // entries to replay (best_mode 1)
    int pad;
};
constexpr int kLogCap = 256;  // accepted swaps remembered between two global bests

}  // namespace mp

// K4: greedy run-order, one persistent CTA.
//
// Replaces order::order_design (src/order.rs:34-105).  The reference re-scores the
// whole prefix for every remaining point at every step (O(n^4 d)).  Appending
// point c to the ordered prefix changes the criterion only through the pairs
// (i, c), i in prefix, so each remaining point carries one running value:
//   maxpro : acc[c] = sum_{i in prefix} 1/(prod_k (x_ik - x_ck)^2 + eps)
//            psi(prefix + c) = ((S_prefix + acc[c]) / pairs)^(1/d) is strictly
//            increasing in acc[c], so the reference's "strictly smaller psi,
//            first position wins" is argmin acc[c] with the same tie rule;
//   maximin: acc[c] = min_{i in prefix} dist(i, c); value(prefix + c) =
//            min(cur_min, acc[c]) -- exact, so ties (which are the common case
//            once cur_min is small) resolve exactly as in the reference.
// Each step adds ONE point, so every acc[c] takes one O(d) update: O(n^2 d)
// total.  The pool keeps Vec::swap_remove order (order.rs:63,101): the last
// element moves into the hole, and scan position decides ties.
// The n-1 steps are strictly sequential (argbest, then update), hence one CTA:
// a grid-wide barrier per step would cost more than the step.
#include "common.cuh"
#include "kernels.h"

namespace mp {

struct OrderParams {
    const double* x;
    unsigned long long n, d;
    int maxpro, minimize;
    unsigned long long* perm;
    double* acc;
    unsigned int* pool;
};

struct Cand {
    double v;
    unsigned int pos;
};

// true if a beats b: strictly better value, or equal value at an earlier scan position
template <bool MINIMIZE>
__device__ __forceinline__ bool cand_better(double va, unsigned int pa, double vb, unsigned int pb) {
    if (va == vb) return pa < pb;
    return MINIMIZE ? (va < vb) : (va > vb);
}

template <bool MAXPRO, bool MINIMIZE>
__global__ void __launch_bounds__(1024, 1) order_kernel(OrderParams p) {
    extern __shared__ __align__(16) double last_s[];  // [d] coordinates of the point just placed
    __shared__ double red_v[32];
    __shared__ unsigned int red_p[32];
    __shared__ unsigned int pick_s;
    __shared__ double cur_min_s;
    const unsigned int n = (unsigned int)p.n, d = (unsigned int)p.d;
    const unsigned int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;

    // first point: closest to the centre [0.5; d], strict <, first index wins (order.rs:49-59)
    double bv = CUDART_INF;
    unsigned int bp = 0xFFFFFFFFu;
    for (unsigned int i = tid; i < n; i += T) {
        const double* xi = p.x + (unsigned long long)i * d;
        double df = __dsub_rn(0.5, xi[0]);
        double sq = __dmul_rn(df, df);
        for (unsigned int k = 1; k < d; ++k) {
            df = __dsub_rn(0.5, xi[k]);
            sq = __dadd_rn(sq, __dmul_rn(df, df));
        }
        double dist = sqrt(sq);
        if (cand_better<true>(dist, i, bv, bp)) { bv = dist; bp = i; }
        p.pool[i] = i;
        p.acc[i] = MAXPRO ? 0.0 : CUDART_INF;
    }
    unsigned int pool_n = n, m = 0;

    auto block_argbest = [&](bool minimize_centre) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            unsigned int op = __shfl_xor_sync(0xffffffffu, bp, off);
            bool take = minimize_centre ? cand_better<true>(ov, op, bv, bp) : cand_better<MINIMIZE>(ov, op, bv, bp);
            if (op != 0xFFFFFFFFu && (bp == 0xFFFFFFFFu || take)) { bv = ov; bp = op; }
        }
        if (lane == 0) { red_v[warp] = bv; red_p[warp] = bp; }
        __syncthreads();
        if (tid == 0) {
            double v = red_v[0];
            unsigned int q = red_p[0];
            for (unsigned int w = 1; w < nw; ++w) {
                if (red_p[w] == 0xFFFFFFFFu) continue;
                bool take = minimize_centre ? cand_better<true>(red_v[w], red_p[w], v, q)
                                            : cand_better<MINIMIZE>(red_v[w], red_p[w], v, q);
                if (q == 0xFFFFFFFFu || take) { v = red_v[w]; q = red_p[w]; }
            }
            // swap_remove(q): the last pool entry fills the hole (order.rs:63,101)
            unsigned int pick = p.pool[q];
            p.pool[q] = p.pool[pool_n - 1];
            p.perm[m] = pick;
            pick_s = pick;
            if (!MAXPRO && !minimize_centre) cur_min_s = fmin(cur_min_s, p.acc[pick]);
        }
        __syncthreads();
    };

    if (tid == 0) cur_min_s = CUDART_INF;
    __syncthreads();
    block_argbest(true);
    --pool_n; ++m;

    while (pool_n > 0) {
        const unsigned int last = pick_s;
        for (unsigned int k = tid; k < d; k += T) last_s[k] = p.x[(unsigned long long)last * d + k];
        const double cur_min = cur_min_s;
        __syncthreads();
        bv = MINIMIZE ? CUDART_INF : -CUDART_INF;
        bp = 0xFFFFFFFFu;
        for (unsigned int pos = tid; pos < pool_n; pos += T) {
            const unsigned int c = p.pool[pos];
            const double* xc = p.x + (unsigned long long)c * d;
            double v;
            if (MAXPRO) {
                double prod = 1.0;  // src/maxpro_utils.rs:45-54, one pair
                for (unsigned int k = 0; k < d; ++k) {
                    double df = __dsub_rn(last_s[k], xc[k]);
                    prod = __dmul_rn(prod, __dmul_rn(df, df));
                }
                v = __dadd_rn(p.acc[c], 1.0 / __dadd_rn(prod, kEps));
                p.acc[c] = v;
            } else {
                double df = __dsub_rn(last_s[0], xc[0]);
                double sq = __dmul_rn(df, df);
                for (unsigned int k = 1; k < d; ++k) {
                    df = __dsub_rn(last_s[k], xc[k]);
                    sq = __dadd_rn(sq, __dmul_rn(df, df));
                }
                double a = fmin(p.acc[c], sqrt(sq));
                p.acc[c] = a;
                v = fmin(a, cur_min);  // criterion of prefix + c
            }
            if (cand_better<MINIMIZE>(v, pos, bv, bp) || bp == 0xFFFFFFFFu) { bv = v; bp = pos; }
        }
        block_argbest(false);
        --pool_n; ++m;
    }
}

cudaError_t launch_order(const OrderLaunch& L) {
    OrderParams p;
    p.x = L.design; p.n = L.n; p.d = L.d; p.maxpro = L.metric == 0; p.minimize = L.minimize;
    p.perm = L.perm; p.acc = L.acc; p.pool = L.pool;
    size_t smem = L.d * sizeof(double);
    int threads = L.n >= 4096 ? 1024 : (L.n >= 1024 ? 512 : 256);
    if (L.metric == 0) {
        if (L.minimize) order_kernel<true, true><<<1, threads, smem, L.stream>>>(p);
        else order_kernel<true, false><<<1, threads, smem, L.stream>>>(p);
    } else {
        if (L.minimize) order_kernel<false, true><<<1, threads, smem, L.stream>>>(p);
        else order_kernel<false, false><<<1, threads, smem, L.stream>>>(p);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

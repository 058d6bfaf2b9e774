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
// Each step adds ONE point, so every running value takes one O(d) update: O(n^2 d)
// total.  The pool keeps Vec::swap_remove order (order.rs:63,101): the last
// element moves into the hole, and scan position decides ties.
// The n-1 steps are strictly sequential (argbest, then update), hence one CTA:
// a grid-wide barrier per step would cost more than the step.
#include "common.cuh"
#include "kernels.h"
#include "order_psi.cuh"

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

// Pool state lives in shared memory, indexed by scan POSITION: pool_s[pos] = point id,
// acc_s[pos] = that point's running value.  swap_remove moves the last position into the hole
// (both arrays), so position order is exactly the reference's Vec order.  Coordinates stay in
// HBM/L2 (row-major, one 8*d-byte row per point): each thread issues the loads of all its
// positions before using any of them.  SMEM: n*12 bytes (+ d doubles), i.e. n <= ~18,000.
constexpr int kOrderMaxPerThread = 8;

template <bool MAXPRO, bool MINIMIZE, bool SMEM_STATE>
__global__ void __launch_bounds__(512, 1) order_kernel(OrderParams p) {
    extern __shared__ __align__(16) unsigned char order_smem[];
    __shared__ double red_v[32];
    __shared__ unsigned int red_p[32];
    __shared__ unsigned int pick_s;
    __shared__ double cur_min_s, best_v_s, s_prefix_s;
    __shared__ unsigned int best_q_s;
    const unsigned int n = (unsigned int)p.n, d = (unsigned int)p.d;
    const unsigned int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    double* last_s = reinterpret_cast<double*>(order_smem);                        // [d]
    double* acc = SMEM_STATE ? last_s + d : p.acc;                                  // [n] by position
    unsigned int* pool = SMEM_STATE ? reinterpret_cast<unsigned int*>(acc + n) : p.pool;  // [n] by position

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
        pool[i] = i;
        acc[i] = MAXPRO ? 0.0 : CUDART_INF;
    }
    unsigned int pool_n = n, m = 0;
    // 16-byte loads need 16-byte aligned rows: even d and an aligned base
    const bool vec_ok = (d % 2 == 0) && ((reinterpret_cast<unsigned long long>(p.x) & 15ull) == 0);

    // block-level argbest of the per-thread (bv, bp): the winning position lands in best_q_s, its value in best_v_s
    auto block_reduce = [&](bool minimize_centre) {
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
            best_v_s = v;
            best_q_s = q;
        }
        __syncthreads();
    };

    auto block_argbest = [&](bool minimize_centre) {
        block_reduce(minimize_centre);
        if (MAXPRO && !minimize_centre) {
            // The reference compares psi of prefix + candidate (order.rs:84-99), first position on equal psi.  Other
            // pool points whose running value lies in the band where psi can coincide with the winner's?  Then
            // the contenders are compared on psi itself (order_psi.cuh).
            const double v = best_v_s, s_prefix = s_prefix_s;
            const unsigned int q0 = best_q_s;
            const double band = order_tie_band(s_prefix, v, d);
            const double lim = MINIMIZE ? __dadd_rn(v, band) : __dsub_rn(v, band);
            int contender = 0;
            for (unsigned int pos = tid; pos < pool_n; pos += T) {
                const double a = acc[pos];
                contender |= (pos != q0 && (MINIMIZE ? a <= lim : a >= lim)) ? 1 : 0;
            }
            if (__syncthreads_or(contender)) {
                const double n_pairs = order_pairs(m + 1), inv_d = 1.0 / (double)d;
                bv = MINIMIZE ? CUDART_INF : -CUDART_INF;
                bp = 0xFFFFFFFFu;
                for (unsigned int pos = tid; pos < pool_n; pos += T) {
                    const double a = acc[pos];
                    if (MINIMIZE ? a <= lim : a >= lim) {
                        const double psi = order_psi(s_prefix, a, n_pairs, inv_d);
                        if (bp == 0xFFFFFFFFu || cand_better<MINIMIZE>(psi, pos, bv, bp)) { bv = psi; bp = pos; }
                    }
                }
                block_reduce(false);
            }
        }
        if (tid == 0) {
            // swap_remove(q): the last pool entry fills the hole (order.rs:63,101)
            const unsigned int q = best_q_s;
            const unsigned int pick = pool[q];
            if (!MAXPRO && !minimize_centre) cur_min_s = fmin(cur_min_s, acc[q]);
            if (MAXPRO && !minimize_centre) s_prefix_s = __dadd_rn(s_prefix_s, acc[q]);  // S of the ordered prefix, in pick order
            pool[q] = pool[pool_n - 1];
            acc[q] = acc[pool_n - 1];
            p.perm[m] = pick;
            pick_s = pick;
        }
        __syncthreads();
    };

    if (tid == 0) { cur_min_s = CUDART_INF; s_prefix_s = 0.0; }
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
        for (unsigned int pos0 = tid; pos0 < pool_n; pos0 += T * kOrderMaxPerThread) {
            // issue every row pointer first, then walk the dimensions of all rows together
            const double* rows[kOrderMaxPerThread];
            double val[kOrderMaxPerThread];
#pragma unroll
            for (int u = 0; u < kOrderMaxPerThread; ++u) {
                const unsigned int pos = pos0 + u * T;
                rows[u] = p.x + (unsigned long long)(pos < pool_n ? pool[pos] : 0u) * d;
                val[u] = MAXPRO ? 1.0 : 0.0;
            }
            auto fold = [&](int u, unsigned int k, double xk) {
                const double df = __dsub_rn(last_s[k], xk);
                if (MAXPRO) val[u] = __dmul_rn(val[u], __dmul_rn(df, df));  // src/maxpro_utils.rs:48-50
                else val[u] = (k == 0) ? __dmul_rn(df, df) : __dadd_rn(val[u], __dmul_rn(df, df));
            };
            // The rows are scattered over HBM/L2 (pool order), so the cost is load latency: fetch
            // four dimensions of every row of the batch (two 16-byte loads per row, all in
            // flight together) before touching any of them.
            unsigned int k = 0;
            if (vec_ok) {
                for (; k + 4 <= d; k += 4) {
                    double2 a[kOrderMaxPerThread], b[kOrderMaxPerThread];
#pragma unroll
                    for (int u = 0; u < kOrderMaxPerThread; ++u) {
                        a[u] = __ldg(reinterpret_cast<const double2*>(rows[u] + k));
                        b[u] = __ldg(reinterpret_cast<const double2*>(rows[u] + k + 2));
                    }
#pragma unroll
                    for (int u = 0; u < kOrderMaxPerThread; ++u) {
                        if (pos0 + u * T < pool_n) {
                            fold(u, k, a[u].x); fold(u, k + 1, a[u].y); fold(u, k + 2, b[u].x); fold(u, k + 3, b[u].y);
                        }
                    }
                }
            }
            for (; k < d; ++k) {
                double xk[kOrderMaxPerThread];
#pragma unroll
                for (int u = 0; u < kOrderMaxPerThread; ++u) xk[u] = __ldg(rows[u] + k);
#pragma unroll
                for (int u = 0; u < kOrderMaxPerThread; ++u)
                    if (pos0 + u * T < pool_n) fold(u, k, xk[u]);
            }
#pragma unroll
            for (int u = 0; u < kOrderMaxPerThread; ++u) {
                const unsigned int pos = pos0 + u * T;
                if (pos < pool_n) {
                    double v;
                    if (MAXPRO) {
                        v = __dadd_rn(acc[pos], 1.0 / __dadd_rn(val[u], kEps));  // :53-54, one pair
                        acc[pos] = v;
                    } else {
                        const double a = fmin(acc[pos], sqrt(val[u]));
                        acc[pos] = a;
                        v = fmin(a, cur_min);  // criterion of prefix + this point
                    }
                    if (cand_better<MINIMIZE>(v, pos, bv, bp) || bp == 0xFFFFFFFFu) { bv = v; bp = pos; }
                }
            }
        }
        block_argbest(false);
        --pool_n; ++m;
    }
}

cudaError_t launch_order(const OrderLaunch& L) {
    OrderParams p;
    p.x = L.design; p.n = L.n; p.d = L.d; p.maxpro = L.metric == 0; p.minimize = L.minimize;
    p.perm = L.perm; p.acc = L.acc; p.pool = L.pool;
    size_t smem_state = L.d * sizeof(double) + L.n * (sizeof(double) + sizeof(unsigned int));
    const bool in_smem = smem_state + 1024 <= kSmemBudget;
    size_t smem = in_smem ? smem_state : L.d * sizeof(double);
    int threads = L.n >= 1024 ? 512 : 256;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<1, threads, smem, L.stream>>>(p);
        count_launch();
        return cudaGetLastError();
    };
    if (L.metric == 0) {
        if (in_smem) return L.minimize ? go(order_kernel<true, true, true>) : go(order_kernel<true, false, true>);
        return L.minimize ? go(order_kernel<true, true, false>) : go(order_kernel<true, false, false>);
    }
    if (in_smem) return L.minimize ? go(order_kernel<false, true, true>) : go(order_kernel<false, false, true>);
    return L.minimize ? go(order_kernel<false, true, false>) : go(order_kernel<false, false, false>);
}

}  // namespace mp

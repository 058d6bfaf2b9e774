// K3: many independent Metropolis chains, one CTA per chain.
//
// Replaces anneal::anneal_lhd + swap_rows (src/anneal.rs:23-143).  The reference
// runs ONE chain and re-scores the whole design (O(n^2 d)) every iteration; here
//   - each chain keeps its current design in shared memory, dimension-major
//     [k][i], for the whole run (160 KB at n=1000, d=20: one chain per SM);
//   - swap moves on the MaxPro criterion are scored incrementally: exchanging
//     x[r1][c] and x[r2][c] only changes the 2(n-2) pairs that touch r1 or r2,
//     and each change is evaluated as (p - p')/(p p') so the update does not
//     cancel; S is re-summed from scratch every kResync iterations to stop drift;
//   - jitter moves (every coordinate changes) and maximin moves are re-scored in
//     full, with the row pairs dealt cyclically (i, i+s mod n) so every thread
//     gets the same number of pairs;
//   - the chain's global best (anneal.rs:92-94,131-136) lives in HBM and is
//     touched only when it improves; when the current design was already the
//     best, a swap improvement rewrites two elements instead of the design.
// Control flow per iteration is the reference's: move, score, diff (negated when
// maximising), p = diff < 0 ? 1 : exp(-diff/temp), one uniform draw ALWAYS
// consumed, accept if u < p, strict-improvement global best, temp *= cooling.
// Draws: Philox key = chain stream (seed + chain), ctr = {t_lo, t_hi, block, domain}.
#include "common.cuh"
#include "philox.cuh"
#include "kernels.h"

namespace mp {

struct AnnealParams {
    const double* design;  // [n][d] row-major start design (shared by all chains)
    unsigned long long n, d, iters, seed, n_chains;
    double t0, cooling, step;
    int minimize, swap;
    unsigned long long resync;
    double n_pairs, inv_d;
    double* best_design;        // [n_chains][n*d] row-major, valid where improved[c] != 0
    double* best_metric;        // [n_chains]
    unsigned int* improved;     // [n_chains]
};

__device__ __forceinline__ double block_reduce(double v, double* scratch, bool is_sum) {
    // deterministic: fixed lane tree, then warp partials folded in warp order
    v = is_sum ? warp_reduce_sum(v) : warp_reduce_min(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = scratch[0];
    for (int w = 1; w < nw; ++w) r = is_sum ? r + scratch[w] : fmin(r, scratch[w]);
    return r;
}

// Full criterion of the design at xs ([k][i] layout).  MAXPRO: S (maxpro_utils.rs:38-58);
// else min squared distance, bit-exact per pair (maximin_utils.rs:39-44).
template <bool MAXPRO>
__device__ __forceinline__ double block_full_score(const double* xs, int n, int d, double* scratch) {
    double t = MAXPRO ? 0.0 : CUDART_INF;
    const int half = (n - 1) / 2;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        // cyclic pairing: row i meets rows i+1 .. i+half (mod n); for even n the
        // diameter pairs (i, i+n/2) are taken by the lower half only
        const int smax = half + ((n % 2 == 0 && i < n / 2) ? 1 : 0);
        for (int s = 1; s <= smax; ++s) {
            int j = i + s;
            if (j >= n) j -= n;
            if (MAXPRO) {
                double p = 1.0;
                for (int k = 0; k < d; ++k) {
                    double df = xs[k * n + i] - xs[k * n + j];
                    p *= df * df;
                }
                t += 1.0 / (p + kEps);
            } else {
                double df = __dsub_rn(xs[i], xs[j]);
                double sq = __dmul_rn(df, df);
                for (int k = 1; k < d; ++k) {
                    df = __dsub_rn(xs[k * n + i], xs[k * n + j]);
                    sq = __dadd_rn(sq, __dmul_rn(df, df));
                }
                t = fmin(t, sq);
            }
        }
    }
    return block_reduce(t, scratch, MAXPRO);
}

template <bool MAXPRO>
__device__ __forceinline__ double metric_from_raw(double raw, int n, double n_pairs, double inv_d) {
    if (MAXPRO) return n < 2 ? 0.0 : pow(raw / n_pairs, inv_d);  // maxpro_utils.rs:75-82
    return sqrt(raw);                                              // maximin_utils.rs:44
}

template <bool MAXPRO>
__global__ void __launch_bounds__(512) anneal_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    double* cur = smem_d;                          // [k][i]
    double* cand = p.swap ? cur : smem_d + nd;     // jitter needs the proposal beside the current
    double* scratch = smem_d + (p.swap ? nd : 2 * nd);  // 32 doubles of reduce scratch

    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            int i = e / d, k = e - i * d;
            cur[k * n + i] = p.design[e];  // anneal.rs:89
        }
        __syncthreads();
        double raw_cur = block_full_score<MAXPRO>(cur, n, d, scratch);
        double cur_metric = metric_from_raw<MAXPRO>(raw_cur, n, p.n_pairs, p.inv_d);  // :90
        double global_best = cur_metric;                                                // :94
        bool materialized = false, synced = true;  // HBM copy exists / equals the current design
        double temp = p.t0;                                                             // :81
        double* best_out = p.best_design + chain * (unsigned long long)nd;
        unsigned long long until_resync = p.resync;  // iterations until the next full re-sum of S

        for (unsigned long long t = 0; t < p.iters; ++t) {
            double raw_new, u;
            int r1 = 0, r2 = 0, c = 0;
            if (p.swap) {
                Philox4 w = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0u, kDomainSwap, key0, key1);
                r1 = (int)mulhi_bound(w.x, (uint32_t)n);  // anneal.rs:29
                r2 = (int)mulhi_bound(w.y, (uint32_t)n);  // :30
                c = (int)mulhi_bound(w.z, (uint32_t)d);   // :31
                u = (double)w.w * 0x1p-32;
                if (MAXPRO && --until_resync != 0) {
                    // incremental: only pairs (r1,j), (r2,j), j not in {r1,r2}, change
                    double delta = 0.0;
                    if (r1 != r2) {
                        const double a_c = cur[c * n + r1], b_c = cur[c * n + r2];
                        for (int j = threadIdx.x; j < n; j += blockDim.x) {
                            if (j == r1 || j == r2) continue;
                            double A = 1.0, B = 1.0;
                            for (int k = 0; k < d; ++k) {
                                if (k == c) continue;
                                const double xj = cur[k * n + j];
                                A *= cur[k * n + r1] - xj;
                                B *= cur[k * n + r2] - xj;
                            }
                            const double xjc = cur[c * n + j];
                            const double uu = a_c - xjc, vv = b_c - xjc;
                            const double au = A * uu, av = A * vv, bu = B * uu, bv = B * vv;
                            const double p1 = fma(au, au, kEps), p1n = fma(av, av, kEps);
                            const double p2 = fma(bv, bv, kEps), p2n = fma(bu, bu, kEps);
                            // 1/p' - 1/p = (p - p')/(p p')
                            delta += (p1 - p1n) / (p1 * p1n) + (p2 - p2n) / (p2 * p2n);
                        }
                    }
                    delta = block_reduce(delta, scratch, true);
                    raw_new = raw_cur + delta;
                } else {
                    until_resync = p.resync;
                    // maximin, or the periodic full re-sum: apply, score, (maybe) undo
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        double v1 = cur[c * n + r1], v2 = cur[c * n + r2];
                        cur[c * n + r1] = v2; cur[c * n + r2] = v1;  // anneal.rs:34-39
                    }
                    __syncthreads();
                    raw_new = block_full_score<MAXPRO>(cur, n, d, scratch);
                    __syncthreads();
                    if (threadIdx.x == 0) {  // undo; re-applied below on acceptance
                        double v1 = cur[c * n + r1], v2 = cur[c * n + r2];
                        cur[c * n + r1] = v2; cur[c * n + r2] = v1;
                    }
                    __syncthreads();
                }
            } else {
                // jitter: element e = i*d + k (row-major, the reference's iteration order,
                // anneal.rs:102-107) takes word e%4 of block e/4
                const double two_step = __dadd_rn(p.step, p.step);
                for (int blk = threadIdx.x; blk * 4 < nd; blk += blockDim.x) {
                    Philox4 w = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)blk, kDomainJit, key0, key1);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int e = blk * 4 + q;
                        if (e < nd) {
                            int i = e / d, k = e - i * d;
                            double uu = (double)ws[q] * 0x1p-32;
                            double delta = __dsub_rn(__dmul_rn(uu, two_step), p.step);
                            double v = __dadd_rn(cur[k * n + i], delta);
                            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);  // clamp(0.0, 1.0)
                            cand[k * n + i] = v;
                        }
                    }
                }
                Philox4 wa = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0u, kDomainJacc, key0, key1);
                u = (double)wa.x * 0x1p-32;
                __syncthreads();
                raw_new = block_full_score<MAXPRO>(cand, n, d, scratch);
            }

            // Metropolis (anneal.rs:110-128); every thread evaluates the same numbers
            const double new_metric = metric_from_raw<MAXPRO>(raw_new, n, p.n_pairs, p.inv_d);
            double diff = new_metric - cur_metric;
            if (!p.minimize) diff *= -1.0;
            const double p_acc = diff < 0.0 ? 1.0 : exp(-diff / temp);
            const bool accept = u < p_acc;
            __syncthreads();
            if (accept) {
                if (p.swap) {
                    if (threadIdx.x == 0) {
                        double v1 = cur[c * n + r1], v2 = cur[c * n + r2];
                        cur[c * n + r1] = v2; cur[c * n + r2] = v1;
                    }
                } else {
                    for (int e = threadIdx.x; e < nd; e += blockDim.x) cur[e] = cand[e];
                }
                cur_metric = new_metric;
                raw_cur = raw_new;
                __syncthreads();
            }
            // strict-improvement global best (anneal.rs:131-136)
            const bool improves = (p.minimize && cur_metric < global_best) || (!p.minimize && cur_metric > global_best);
            if (improves) {
                if (p.swap && materialized && synced && accept) {
                    if (threadIdx.x == 0) {
                        best_out[r1 * d + c] = cur[c * n + r1];
                        best_out[r2 * d + c] = cur[c * n + r2];
                    }
                } else {
                    for (int e = threadIdx.x; e < nd; e += blockDim.x) {
                        int i = e / d, k = e - i * d;
                        best_out[e] = cur[k * n + i];
                    }
                }
                global_best = cur_metric;
                materialized = true;
                synced = true;
            } else if (accept) {
                synced = false;
            }
            temp *= p.cooling;  // anneal.rs:139
        }
        if (threadIdx.x == 0) {
            p.best_metric[chain] = global_best;
            p.improved[chain] = materialized ? 1u : 0u;
        }
    }
}

size_t anneal_smem_bytes(unsigned long long n, unsigned long long d, int swap) {
    return (size_t)(n * d * (swap ? 1 : 2) + 32 + 2) * sizeof(double);
}

bool anneal_supported(unsigned long long n, unsigned long long d, int swap) {
    return n >= 1 && d >= 1 && n < (1ull << 31) / (d ? d : 1) && anneal_smem_bytes(n, d, swap) <= kSmemBudget;
}

cudaError_t launch_anneal(const AnnealLaunch& L) {
    AnnealParams p;
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.01 / (double)L.n;  // anneal.rs:86-87
    p.minimize = L.minimize; p.swap = L.swap;
    unsigned long long rs = 16 * L.n;
    p.resync = rs < 1024 ? 1024 : rs;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved;
    size_t smem = anneal_smem_bytes(L.n, L.d, L.swap);
    int threads = L.n <= 64 ? 64 : (L.n <= 128 ? 128 : (L.n <= 512 ? 256 : 512));
    unsigned long long grid = L.n_chains < (1ull << 20) ? L.n_chains : (1ull << 20);
    cudaError_t e;
    if (L.metric == 0) {
        e = cudaFuncSetAttribute(anneal_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        anneal_kernel<true><<<(unsigned int)grid, threads, smem, L.stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(anneal_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        anneal_kernel<false><<<(unsigned int)grid, threads, smem, L.stream>>>(p);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

// K4 on a thread-block CLUSTER: greedy run order with the pool distributed over the shared
// memory of 8 SMs (distributed shared memory, DSMEM).
//
// Same algorithm and tie rules as order.cu (src/order.rs:34-105 with one running value per
// remaining point).  order.cu runs on one SM: every step re-reads the coordinates of all
// remaining points through that SM's L2 port (160 KB per step at n=5000, d=8) and does all the
// FP64 work on 64 lanes -- 7 us per step.  Here the pool is dealt round-robin over the CTAs of a
// cluster: scan position p lives in CTA p mod 8, slot p / 8, with its point id, running value
// AND coordinates in that CTA's shared memory (47 KB per CTA at n=5000, d=8), so a step reads
// nothing from L2/HBM at all.
//
// One step = two cluster barriers:
//   A  every CTA updates its slots against the point just placed, finds its best (value,
//      position), and thread 0 stores that pair into the candidate table of EVERY CTA (DSMEM);
//      -- cluster.sync --
//   B  every CTA folds the 8 candidates (identical result everywhere): position q wins.
//      The owner of q records perm[m], and sends the winner's coordinates (and, for maximin,
//      its running minimum) to every CTA's `last` buffer; the owner of the last position L
//      sends its entry to a one-entry mailbox in the owner of q (Vec::swap_remove moves the last
//      element into the hole, order.rs:63,101);
//      -- cluster.sync --
//   then the owner of q drains its mailbox into slot q/8 and the next step starts.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "order_psi.cuh"

namespace cg = cooperative_groups;

namespace mp {

constexpr int kOrdCluster = 8;
constexpr int kOrdThreads = 256;

struct OrdClParams {
    const double* x;
    unsigned int n, d, ds, slots;  // ds = odd slot stride in doubles (bank-conflict-free), slots per CTA = ceil(n / 8)
    unsigned long long* perm;
};

template <bool MINIMIZE>
__device__ __forceinline__ bool ocl_better(double va, unsigned int pa, double vb, unsigned int pb) {
    if (pb == 0xFFFFFFFFu) return pa != 0xFFFFFFFFu;
    if (pa == 0xFFFFFFFFu) return false;
    if (va == vb) return pa < pb;  // equal value: earlier scan position (order.rs:89-99 strict compare)
    return MINIMIZE ? (va < vb) : (va > vb);
}

template <bool MAXPRO, bool MINIMIZE>
__global__ void __cluster_dims__(kOrdCluster, 1, 1) __launch_bounds__(kOrdThreads, 1) order_cluster_kernel(OrdClParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank();
    const unsigned int n = p.n, d = p.d, ds = p.ds, slots = p.slots;
    const unsigned int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned int T = kOrdThreads, NW = kOrdThreads / 32;

    extern __shared__ __align__(16) unsigned char ocl_smem[];
    double* coords = reinterpret_cast<double*>(ocl_smem);            // [slots][ds]
    double* acc = coords + (size_t)slots * ds;                        // [slots]
    double* last_s = acc + slots;                                     // [d + 1]  (+1: winner's running value)
    double* mail_d = last_s + d + 1;                                  // [d + 1]  mailbox: coords + running value
    double* cand_v = mail_d + d + 1;                                  // [8]
    double* red_v = cand_v + kOrdCluster;                             // [NW]
    unsigned int* ids = reinterpret_cast<unsigned int*>(red_v + NW);  // [slots]
    unsigned int* cand_p = ids + slots;                               // [8]
    unsigned int* red_p = cand_p + kOrdCluster;                       // [NW]
    unsigned int* mail_id = red_p + NW;                               // [1]
    unsigned int* cand_s = mail_id + 1;                               // [8]   second-best key (high word) of each CTA
    unsigned int* red_s = cand_s + kOrdCluster;                       // [NW]
    double* cand2_v = reinterpret_cast<double*>(ocl_smem + ((reinterpret_cast<unsigned char*>(red_s + NW) - ocl_smem + 7) / 8) * 8);  // [8]  psi round
    unsigned int* cand2_p = reinterpret_cast<unsigned int*>(cand2_v + kOrdCluster);  // [8]

    // ---- load this CTA's slots: position p = slot*8 + rank holds point p at the start
    for (unsigned int s = tid; s < slots; s += T) {
        const unsigned int pos = s * kOrdCluster + rank;
        ids[s] = pos;
        acc[s] = MAXPRO ? 0.0 : CUDART_INF;
    }
    for (unsigned int e = tid; e < slots * d; e += T) {
        const unsigned int s = e / d, k = e - s * d, pos = s * kOrdCluster + rank;
        coords[(size_t)s * ds + k] = pos < n ? p.x[(unsigned long long)pos * d + k] : 0.0;
    }
    __syncthreads();

    unsigned int pool_n = n, m = 0;
    double cur_min = CUDART_INF;
    double s_prefix = 0.0;  // MaxPro sum of the ordered prefix, in pick order (every thread keeps its copy)

    // block-level argbest of (bv, bp) -> broadcast to every CTA's candidate table (tv, tp).  s2 = high key word of
    // the thread's second-best value (order_psi.cuh); the CTA's second-best goes to the table ts when it is given.
    auto publish_local_best = [&](double bv, unsigned int bp, unsigned int s2, bool centre, double* tv, unsigned int* tp, unsigned int* ts) {
        double bv0 = bv;
        unsigned int bp0 = bp;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const unsigned int op = __shfl_xor_sync(0xffffffffu, bp, off);
            if (centre ? ocl_better<true>(ov, op, bv, bp) : ocl_better<MINIMIZE>(ov, op, bv, bp)) { bv = ov; bp = op; }
        }
        // a lane whose own best lost the warp's argbest offers it as a second-best
        s2 = __reduce_min_sync(0xffffffffu, min(s2, (bp0 != 0xFFFFFFFFu && bp0 != bp) ? order_key_hi<MINIMIZE>(bv0) : kNoKeyHi));
        if (lane == 0) { red_v[warp] = bv; red_p[warp] = bp; red_s[warp] = s2; }
        __syncthreads();
        if (warp == 0) {
            bv = lane < NW ? red_v[lane] : 0.0;
            bp = lane < NW ? red_p[lane] : 0xFFFFFFFFu;
            s2 = lane < NW ? red_s[lane] : kNoKeyHi;
            bv0 = bv; bp0 = bp;
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const unsigned int op = __shfl_xor_sync(0xffffffffu, bp, off);
                if (centre ? ocl_better<true>(ov, op, bv, bp) : ocl_better<MINIMIZE>(ov, op, bv, bp)) { bv = ov; bp = op; }
            }
            const unsigned int bpf = __shfl_sync(0xffffffffu, bp, 0);  // lanes 0..7 agree; the others never held a candidate
            s2 = __reduce_min_sync(0xffffffffu, min(s2, (bp0 != 0xFFFFFFFFu && bp0 != bpf) ? order_key_hi<MINIMIZE>(bv0) : kNoKeyHi));
            if (lane < kOrdCluster) {  // lane r stores into CTA r
                double* rv = cluster.map_shared_rank(tv, lane);
                unsigned int* rp = cluster.map_shared_rank(tp, lane);
                rv[rank] = bv;
                rp[rank] = bp;
                if (ts) cluster.map_shared_rank(ts, lane)[rank] = s2;
            }
        }
    };

    // fold the 8 candidates (same answer in every CTA), then hand-overs of phase B
    auto settle = [&](bool centre) {
        double bv = cand_v[0];
        unsigned int q = cand_p[0];
#pragma unroll
        for (int r = 1; r < kOrdCluster; ++r)
            if (centre ? ocl_better<true>(cand_v[r], cand_p[r], bv, q) : ocl_better<MINIMIZE>(cand_v[r], cand_p[r], bv, q)) {
                bv = cand_v[r]; q = cand_p[r];
            }
        if (MAXPRO && !centre) {
            // The reference compares psi of prefix + candidate (order.rs:84-99), first position on equal psi.  Is there
            // another pool point whose running value lies in the band where psi can coincide with the winner's?
            // Every CTA reads the same table, so all of them take the branch together.
            unsigned int second = kNoKeyHi;
#pragma unroll
            for (int r = 0; r < kOrdCluster; ++r) {
                if (cand_p[r] == 0xFFFFFFFFu) continue;
                second = min(second, cand_s[r]);
                if (cand_p[r] != q) second = min(second, order_key_hi<MINIMIZE>(cand_v[r]));
            }
            const double band = order_tie_band(s_prefix, bv, d);
            const double lim = MINIMIZE ? __dadd_rn(bv, band) : __dsub_rn(bv, band);
            if (second <= order_key_hi<MINIMIZE>(lim)) {
                const double n_pairs = order_pairs(m + 1), inv_d = 1.0 / (double)d;
                double pv = MINIMIZE ? CUDART_INF : -CUDART_INF;
                unsigned int pp = 0xFFFFFFFFu;
                for (unsigned int s = tid; s < slots; s += T) {
                    const unsigned int pos = s * kOrdCluster + rank;
                    if (pos >= pool_n) break;
                    const double a = acc[s];
                    if (MINIMIZE ? a <= lim : a >= lim) {
                        const double psi = order_psi(s_prefix, a, n_pairs, inv_d);
                        if (ocl_better<MINIMIZE>(psi, pos, pv, pp)) { pv = psi; pp = pos; }
                    }
                }
                publish_local_best(pv, pp, kNoKeyHi, false, cand2_v, cand2_p, nullptr);
                cluster.sync();
                q = cand2_p[0];
                bv = cand2_v[0];
#pragma unroll
                for (int r = 1; r < kOrdCluster; ++r)
                    if (ocl_better<MINIMIZE>(cand2_v[r], cand2_p[r], bv, q)) { bv = cand2_v[r]; q = cand2_p[r]; }
            }
        }
        const unsigned int L = pool_n - 1;
        const unsigned int q_owner = q % kOrdCluster, q_slot = q / kOrdCluster;
        const unsigned int l_owner = L % kOrdCluster, l_slot = L / kOrdCluster;
        if (rank == q_owner) {
            if (tid == 0) p.perm[m] = ids[q_slot];
            // winner's coordinates (+ running value) to every CTA
            for (unsigned int e = tid; e < (d + 1) * kOrdCluster; e += T) {
                const unsigned int r = e / (d + 1), k = e - r * (d + 1);
                double* dst = cluster.map_shared_rank(last_s, r);
                dst[k] = k < d ? coords[(size_t)q_slot * ds + k] : acc[q_slot];
            }
        }
        if (rank == l_owner && L != q) {  // the last entry moves into the hole: mail it to q's owner
            double* md = cluster.map_shared_rank(mail_d, q_owner);
            unsigned int* mi = cluster.map_shared_rank(mail_id, q_owner);
            for (unsigned int k = tid; k <= d; k += T) md[k] = k < d ? coords[(size_t)l_slot * ds + k] : acc[l_slot];
            if (tid == 0) mi[0] = ids[l_slot];
        }
        cluster.sync();
        if (rank == q_owner && L != q) {  // drain the mailbox into the hole
            for (unsigned int k = tid; k <= d; k += T) {
                if (k < d) coords[(size_t)q_slot * ds + k] = mail_d[k];
                else acc[q_slot] = mail_d[d];
            }
            if (tid == 0) ids[q_slot] = mail_id[0];
        }
        if (!MAXPRO && !centre) cur_min = fmin(cur_min, last_s[d]);
        if (MAXPRO && !centre) s_prefix = __dadd_rn(s_prefix, last_s[d]);
        __syncthreads();
        --pool_n; ++m;
    };

    // ---- first point: closest to the centre, strict <, first index wins (order.rs:49-59)
    {
        double bv = CUDART_INF;
        unsigned int bp = 0xFFFFFFFFu;
        for (unsigned int s = tid; s < slots; s += T) {
            const unsigned int pos = s * kOrdCluster + rank;
            if (pos >= n) continue;
            const double* xi = coords + (size_t)s * ds;
            double df = __dsub_rn(0.5, xi[0]);
            double sq = __dmul_rn(df, df);
            for (unsigned int k = 1; k < d; ++k) {
                df = __dsub_rn(0.5, xi[k]);
                sq = __dadd_rn(sq, __dmul_rn(df, df));
            }
            const double dist = sqrt(sq);
            if (ocl_better<true>(dist, pos, bv, bp)) { bv = dist; bp = pos; }
        }
        publish_local_best(bv, bp, kNoKeyHi, true, cand_v, cand_p, cand_s);
        cluster.sync();
        settle(true);
    }

    // ---- greedy steps
    while (pool_n > 0) {
        double bv = MINIMIZE ? CUDART_INF : -CUDART_INF;
        unsigned int bp = 0xFFFFFFFFu, s2 = kNoKeyHi;
        for (unsigned int s = tid; s < slots; s += T) {
            const unsigned int pos = s * kOrdCluster + rank;
            if (pos >= pool_n) break;
            const double* xc = coords + (size_t)s * ds;
            double v;
            if (MAXPRO) {
                double prod = 1.0;  // src/maxpro_utils.rs:45-54, one pair
                for (unsigned int k = 0; k < d; ++k) {
                    const double df = __dsub_rn(last_s[k], xc[k]);
                    prod = __dmul_rn(prod, __dmul_rn(df, df));
                }
                v = __dadd_rn(acc[s], 1.0 / __dadd_rn(prod, kEps));
                acc[s] = v;
            } else {
                double df = __dsub_rn(last_s[0], xc[0]);
                double sq = __dmul_rn(df, df);
                for (unsigned int k = 1; k < d; ++k) {
                    df = __dsub_rn(last_s[k], xc[k]);
                    sq = __dadd_rn(sq, __dmul_rn(df, df));
                }
                const double a = fmin(acc[s], sqrt(sq));
                acc[s] = a;
                v = fmin(a, cur_min);  // criterion of prefix + this point
            }
            if (ocl_better<MINIMIZE>(v, pos, bv, bp)) {
                if (MAXPRO && bp != 0xFFFFFFFFu) s2 = min(s2, order_key_hi<MINIMIZE>(bv));
                bv = v; bp = pos;
            } else if (MAXPRO) {
                s2 = min(s2, order_key_hi<MINIMIZE>(v));
            }
        }
        publish_local_best(bv, bp, s2, false, cand_v, cand_p, cand_s);
        cluster.sync();
        settle(false);
    }
}

static size_t order_cluster_smem(unsigned long long n, unsigned long long d) {
    unsigned long long slots = (n + kOrdCluster - 1) / kOrdCluster;
    return (size_t)(slots * (d | 1ull) * 8 + slots * 8 + 2 * (d + 1) * 8 + kOrdCluster * 8 + (kOrdThreads / 32) * 8 +
                    slots * 4 + kOrdCluster * 4 + (kOrdThreads / 32) * 4 + 16 +
                    kOrdCluster * 4 + (kOrdThreads / 32) * 4 + 8 + kOrdCluster * 12);  // second keys + the psi round's table
}

bool order_cluster_supported(unsigned long long n, unsigned long long d) {
    return n >= 64 && n < (1ull << 31) && d >= 1 && order_cluster_smem(n, d) + 1024 <= kSmemBudget;
}

cudaError_t launch_order_cluster(const OrderLaunch& L) {
    OrdClParams p;
    p.x = L.design; p.n = (unsigned int)L.n; p.d = (unsigned int)L.d; p.ds = (unsigned int)(L.d | 1ull);
    p.slots = (unsigned int)((L.n + kOrdCluster - 1) / kOrdCluster);
    p.perm = L.perm;
    const size_t smem = order_cluster_smem(L.n, L.d);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<kOrdCluster, kOrdThreads, smem, L.stream>>>(p);  // one cluster (compile-time __cluster_dims__)
        count_launch();
        return cudaGetLastError();
    };
    if (L.metric == 0) return L.minimize ? go(order_cluster_kernel<true, true>) : go(order_cluster_kernel<true, false>);
    return L.minimize ? go(order_cluster_kernel<false, true>) : go(order_cluster_kernel<false, false>);
}

}  // namespace mp

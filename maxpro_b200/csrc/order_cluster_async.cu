// K4 on a thread-block cluster WITHOUT cluster barriers in the step: the exchanges of
// order_cluster.cu done with mbarriers, st.async (remote store that completes bytes on the
// destination's mbarrier) and remote loads through distributed shared memory.
//
// Same algorithm, pool layout and tie rules as order_cluster.cu (src/order.rs:34-105 with one
// running value per remaining point; scan position p lives in CTA p mod 8, slot p / 8).  That
// kernel pays two `cluster.sync` per greedy step (~380 cycles each, every thread of 8 CTAs) plus a
// remote store in front of each.  Here a step is
//
//   1  the worker warps update the CTA's slots against the point just placed; every WARP finds its best
//      (value, position) with three REDUX.MIN on integer keys and sends that 16-byte record to the table of
//      every CTA of the cluster with st.async, which counts the bytes on the DESTINATION's mbarrier; each
//      CTA waits on its own mbarrier only (tables and mbarriers double-buffered by step parity, every phase
//      armed in advance).  A complete table also says that every warp of the cluster is past its slots;
//   2  every warp folds the records (same answer everywhere: position q wins) and warp 0 PULLS the winner's
//      coordinates (and running value) out of its owner's shared memory with d + 1 remote loads; one
//      __syncthreads makes them the CTA's;
//   3  off the critical path: the owner of the last position L mails its entry to the owner of q
//      (Vec::swap_remove moves the last element into the hole, order.rs:63,101) with st.async on the
//      owner's `hole` mbarrier, and every CTA arrives on that mbarrier once its pull has landed.  Warp 0 of
//      the owner, which has no slots of its own, fills the hole in the NEXT step when the mbarrier completes
//      (mail in, nobody still reading the slot) and evaluates the entry straight out of the mailbox.
//
// Critical path per step: slot update, warp argbest, one DSMEM store latency, fold, one DSMEM load latency,
// one CTA barrier.  No thread waits for a thread of another CTA except through those two transfers.
// Pools above 3,584 points take a 16-CTA cluster (non-portable size, launch attribute), the others 8.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "order_psi.cuh"

namespace cg = cooperative_groups;

namespace mp {

namespace {

constexpr unsigned int kNone = 0xFFFFFFFFu;

struct __align__(16) CandRec {
    double v;
    unsigned long long p;       // scan position, kNone = no candidate
};

struct OrdAsParams {
    const double* x;
    unsigned int n, d, ds, slots;  // ds = odd slot stride in doubles, slots per CTA = ceil(n / 8)
    unsigned long long* perm;
};

__device__ __forceinline__ unsigned int smem_u32(const void* ptr) { return (unsigned int)__cvta_generic_to_shared(ptr); }

// address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ unsigned int map_rank(unsigned int addr, unsigned int rank) {
    unsigned int out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
    return out;
}

__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// one arrival of this thread + `bytes` more transaction bytes to wait for
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// arrival on an mbarrier of another CTA; release at cluster scope: what this thread did (and saw) is done
__device__ __forceinline__ void mbar_arrive_remote(unsigned int remote_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}

// Wait for the phase with the given parity.  try_wait sleeps in hardware for a bounded time; a protocol
// error would spin forever, so after two seconds the loop gives up with a trap (the launch then fails
// instead of hanging the device).
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    unsigned int done = 0;
    unsigned long long t0 = 0;
    for (unsigned int spin = 1; !done; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if ((spin & 0x3FFu) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}

__device__ __forceinline__ void st_async_v2(unsigned int remote_addr, unsigned long long a, unsigned long long b,
                                            unsigned int remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
                 "l"(a), "l"(b), "r"(remote_bar)
                 : "memory");
}

__device__ __forceinline__ void st_async_b64(unsigned int remote_addr, unsigned long long a, unsigned int remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(a),
                 "r"(remote_bar)
                 : "memory");
}

__device__ __forceinline__ unsigned long long ld_cluster_b64(unsigned int remote_addr) {
    unsigned long long v;
    asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(v) : "r"(remote_addr) : "memory");
    return v;
}

// ---- argbest on integer keys -------------------------------------------------------------------------
// Every value compared here is a non-negative double (a sum of positive terms, a distance, or +inf), so
// the bit pattern orders like the number.  key = bits when the smaller value wins, ~bits when the larger
// one does: the winner is always the SMALLEST (key, position) pair, position breaking ties towards the
// earlier scan position (order.rs:89-99, strict compare).  A NaN never wins in the reference (every
// comparison with it is false); here it gets the worst key.  "No candidate" = (all ones, kNone).
// Integer keys make the reduction three REDUX.MIN per warp (high word, low word among the lanes that
// hold the high minimum, position among the lanes that hold both) instead of five shuffle levels of
// double compares -- measured 610 cycles for the shuffle tree of one warp, 776 for the serial fold of the
// eight cluster records.
template <bool MINIMIZE>
__device__ __forceinline__ unsigned long long value_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (v != v) return ~0ull;
    return MINIMIZE ? b : ~b;
}

__device__ __forceinline__ void warp_argmin_key(unsigned long long& key, unsigned int& pos) {
    const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
    const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    const bool mine = hi == mhi && lo == mlo;
    pos = __reduce_min_sync(0xffffffffu, mine ? pos : kNone);
    key = ((unsigned long long)mhi << 32) | mlo;
}

#ifndef OAS_PROFILE
#define OAS_PROFILE 0  // 1: when each warp of one CTA reaches each point of the step, cycles since the step's barrier (make EXTRA=-DOAS_PROFILE=1)
#endif
#if OAS_PROFILE
#define OAS_T(i) { tsum[i] += clock64() - *t_ref; }  // time since the previous step's barrier, as thread 0 saw it
#else
#define OAS_T(i)
#endif

// THR threads per CTA: warp 0 is the service warp (hole, pull, bookkeeping), the other THR - 32 threads own the slots.  SPP = slots a worker
// evaluates per pass (independent dependency chains side by side).
// CL = CTAs per cluster (launch attribute): 8, or 16 (non-portable size) for pools large enough to pay for the
// doubled record table.
template <bool MAXPRO, bool MINIMIZE, int THR, int SPP, int CL>
__global__ void __launch_bounds__(THR, 1) order_cluster_async_kernel(OrdAsParams p) {
    constexpr unsigned int kCl = CL;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank();
    const unsigned int n = p.n, d = p.d, ds = p.ds, slots = p.slots;
    const unsigned int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned int T = THR, NW = THR / 32, TW = THR - 32;
    constexpr unsigned int NREC = kCl * NW;                 // records per table: one per warp of the cluster
    constexpr unsigned int PER_LANE = (NREC + 31) / 32;

    extern __shared__ __align__(16) unsigned char oas_smem[];
    CandRec* cand = reinterpret_cast<CandRec*>(oas_smem);                     // [2][NREC] candidate tables by step parity
    CandRec* cand_psi = cand + 2 * NREC;                                      // [NREC] table of the psi round (near-ties only)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(cand_psi + NREC);  // [4]  table mbarriers (2), hole mbarrier, pad
    double* coords = reinterpret_cast<double*>(bars + 4);                     // [slots][ds]
    double* acc = coords + (size_t)slots * ds;                                // [slots]
    double* last_s = acc + slots;                                             // [d + 1]  winner: coordinates + running value
    unsigned long long* mail = reinterpret_cast<unsigned long long*>(last_s + d + 1);  // [d + 2]  coordinates, running value, id
    unsigned int* ids = reinterpret_cast<unsigned int*>(mail + d + 2);        // [slots]

    const unsigned int cand_bar0 = smem_u32(&bars[0]), hole_bar = smem_u32(&bars[2]);

    if (tid == 0) {
        mbar_init(cand_bar0, 1);       // thread 0's arrive.expect_tx; the records arrive as transaction bytes
        mbar_init(cand_bar0 + 8, 1);
        mbar_init(hole_bar, kCl);      // one arrival per CTA (pull finished) + the mailed entry as transaction bytes
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // Every phase is ARMED IN ADVANCE with this CTA's own arrival and the bytes it will receive (they never
        // change): arrive.expect_tx costs its thread ~300 cycles, which would otherwise sit in front of a wait.
        mbar_arrive_expect_tx(cand_bar0, NREC * (unsigned int)sizeof(CandRec));
        mbar_arrive_expect_tx(cand_bar0 + 8, NREC * (unsigned int)sizeof(CandRec));
        mbar_arrive_expect_tx(hole_bar, (d + 2) * 8u);
    }
    // ---- load this CTA's slots: position p = slot*8 + rank holds point p at the start
    for (unsigned int s = tid; s < slots; s += T) {
        ids[s] = s * kCl + rank;
        acc[s] = MAXPRO ? 0.0 : CUDART_INF;
    }
    for (unsigned int e = tid; e < slots * d; e += T) {
        const unsigned int s = e / d, k = e - s * d, pos = s * kCl + rank;
        coords[(size_t)s * ds + k] = pos < n ? p.x[(unsigned long long)pos * d + k] : 0.0;
    }
    cluster.sync();  // mbarriers initialised and slots loaded everywhere before any remote access

    unsigned int pool_n = n, m = 0, step = 0, hole_phase = 0, hole_slot = kNone;
    double cur_min = CUDART_INF;
    double s_prefix = 0.0;  // MaxPro sum of the ordered prefix, in pick order (every thread keeps its copy)
#if OAS_PROFILE
    long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __shared__ long long t_ref_s;
    volatile long long* t_ref = &t_ref_s;
    if (tid == 0) t_ref_s = clock64();
    __syncthreads();
#endif

    // One greedy step.  CENTRE = the first point: distance to the centre of the cube, smaller wins, first
    // index on ties (order.rs:49-59); afterwards: the running value against the point just placed (last_s).
    auto greedy_step = [&](auto centre_tag) {
        constexpr bool CENTRE = decltype(centre_tag)::value;
        constexpr bool MINI = CENTRE ? true : MINIMIZE;
        // values of N entries side by side: coordinates x[u], old running values a[u] -> v[u].  Every entry is
        // computed in the reference's order (products and sums walk the dimensions one by one).
        auto eval_n = [&](auto n_tag, const double* const (&x)[SPP], const double (&a)[SPP], double (&v)[SPP]) {
            constexpr int N = decltype(n_tag)::value;
            double r[N];
            if (CENTRE || !MAXPRO) {
#pragma unroll
                for (int u = 0; u < N; ++u) {
                    const double f = __dsub_rn(CENTRE ? 0.5 : last_s[0], x[u][0]);
                    r[u] = __dmul_rn(f, f);
                }
                for (unsigned int k = 1; k < d; ++k) {
                    const double l = CENTRE ? 0.5 : last_s[k];
#pragma unroll
                    for (int u = 0; u < N; ++u) {
                        const double f = __dsub_rn(l, x[u][k]);
                        r[u] = __dadd_rn(r[u], __dmul_rn(f, f));
                    }
                }
#pragma unroll
                for (int u = 0; u < N; ++u) v[u] = CENTRE ? sqrt(r[u]) : fmin(a[u], sqrt(r[u]));
            } else {  // src/maxpro_utils.rs:45-54, one pair
#pragma unroll
                for (int u = 0; u < N; ++u) r[u] = 1.0;
                for (unsigned int k = 0; k < d; ++k) {
                    const double l = last_s[k];
#pragma unroll
                    for (int u = 0; u < N; ++u) {
                        const double f = __dsub_rn(l, x[u][k]);
                        r[u] = __dmul_rn(r[u], __dmul_rn(f, f));
                    }
                }
#pragma unroll
                for (int u = 0; u < N; ++u) v[u] = __dadd_rn(a[u], 1.0 / __dadd_rn(r[u], kEps));
            }
        };
        unsigned long long bk = ~0ull;
        unsigned int bp = kNone;
        unsigned int k2 = kNoKeyHi;  // high key word of this thread's second-best (MaxPro steps: contender detection, order_psi.cuh)
        constexpr bool PSI = MAXPRO && !CENTRE;
        // keep the running value, offer (value, position) to the argbest
        auto offer = [&](unsigned int s, double v) {
            double c = v;
            if (!CENTRE) {
                acc[s] = v;
                if (!MAXPRO) c = fmin(v, cur_min);  // maximin: criterion of prefix + this point
            }
            const unsigned long long k = value_key<MINI>(c);
            const unsigned int pos = s * kCl + rank;
            if (k < bk || (k == bk && pos < bp)) {
                if (PSI && bp != kNone) k2 = min(k2, (unsigned int)(bk >> 32));
                bk = k; bp = pos;
            } else if (PSI) {
                k2 = min(k2, (unsigned int)(k >> 32));
            }
        };

        const unsigned int par = step & 1u;
        const unsigned int cbar = cand_bar0 + 8 * par;
        // ---- 1a. the worker warps update this CTA's slots, up to SPP per pass.  A worker whose pass holds the
        //          hole evaluates the stale entry with the others and drops the result.
        const unsigned int live = pool_n > rank ? (pool_n - rank + kCl - 1) / kCl : 0;  // slots with position < pool_n
        if (tid >= 32) {
            for (unsigned int s0 = tid - 32; s0 < live; s0 += SPP * TW) {
                const double* x[SPP];
                double a[SPP], v[SPP];
                unsigned int cnt = 0;
#pragma unroll
                for (int u = 0; u < SPP; ++u) {
                    const unsigned int s = s0 + u * TW;
                    const bool in = s < live;
                    cnt += in;
                    x[u] = coords + (size_t)(in ? s : s0) * ds;
                    a[u] = CENTRE ? 0.0 : acc[in ? s : s0];
                }
                if (SPP >= 4 && cnt == 4) eval_n(std::integral_constant<int, (SPP >= 4 ? 4 : 1)>{}, x, a, v);
                else if (SPP >= 3 && cnt == 3) eval_n(std::integral_constant<int, (SPP >= 3 ? 3 : 1)>{}, x, a, v);
                else if (SPP >= 2 && cnt == 2) eval_n(std::integral_constant<int, (SPP >= 2 ? 2 : 1)>{}, x, a, v);
                else eval_n(std::integral_constant<int, 1>{}, x, a, v);
#pragma unroll
                for (int u = 0; u < SPP; ++u) {
                    const unsigned int s = s0 + u * TW;
                    if (u < (int)cnt && s != hole_slot) offer(s, v[u]);
                }
            }
        }
        OAS_T(0)
        // ---- 1b. warp 0 has no slots of its own: it fills the hole left by the previous winner once the
        //          mail is in and every CTA has pulled the winner's coordinates out of the slot.  Lane 0 evaluates
        //          the entry straight out of the mailbox while the other lanes copy it into the slot.
        if (hole_slot != kNone) {
            if (tid < 32) {
                mbar_wait(hole_bar, hole_phase);
                double* xc = coords + (size_t)hole_slot * ds;
                for (unsigned int k = lane; k < d; k += 32) xc[k] = __longlong_as_double((long long)mail[k]);
                if (lane == 1) ids[hole_slot] = (unsigned int)mail[d + 1];
                if (lane == 0) {
                    const double* x[SPP];
                    double a[SPP], v[SPP];
#pragma unroll
                    for (int u = 0; u < SPP; ++u) { x[u] = reinterpret_cast<const double*>(mail); a[u] = __longlong_as_double((long long)mail[d]); }
                    eval_n(std::integral_constant<int, 1>{}, x, a, v);
                    offer(hole_slot, v[0]);
                }
                __syncwarp();
            }
            hole_phase ^= 1u;
        }
        // ---- 1c. argbest of the warp; its record goes to every CTA of the cluster (lane r stores into CTA r)
        const unsigned long long own_k = bk;
        const unsigned int own_p = bp;
        warp_argmin_key(bk, bp);
        // second-best of the warp: a lane whose own best lost offers it (high key word only)
        if (PSI) k2 = __reduce_min_sync(0xffffffffu, min(k2, (own_p != kNone && own_p != bp) ? (unsigned int)(own_k >> 32) : kNoKeyHi));
        if (lane < kCl)
            st_async_v2(map_rank(smem_u32(&cand[par * NREC + rank * NW + warp]), lane), bk, ((unsigned long long)k2 << 32) | bp, map_rank(cbar, lane));
        // the hole mbarrier's phase has been consumed: arm the next one (any order with the next mail is fine --
        // a phase cannot complete before this arrival)
        if (hole_slot != kNone && tid == 0) mbar_arrive_expect_tx(hole_bar, (d + 2) * 8u);
        OAS_T(1)
        // ---- 2. wait for the records of all warps of the cluster, fold them (same answer in every warp), pull
        //         the winner.  A complete table also says that every warp of the cluster is past its slots.
        mbar_wait(cbar, (step >> 1) & 1u);
        OAS_T(2)
        bk = ~0ull; bp = kNone; k2 = kNoKeyHi;
#pragma unroll
        for (unsigned int j = 0; j < PER_LANE; ++j) {
            const CandRec rec = cand[par * NREC + (lane + 32 * j) % NREC];
            const unsigned long long k = (unsigned long long)__double_as_longlong(rec.v);
            const unsigned int pos = (unsigned int)rec.p;
            if (PSI) k2 = min(k2, (unsigned int)(rec.p >> 32));
            if (k < bk || (k == bk && pos < bp)) {
                if (PSI && bp != kNone) k2 = min(k2, (unsigned int)(bk >> 32));
                bk = k; bp = pos;
            } else if (PSI && pos != kNone && pos != bp) {
                k2 = min(k2, (unsigned int)(k >> 32));
            }
        }
        const unsigned long long own_fk = bk;
        const unsigned int own_fp = bp;
        warp_argmin_key(bk, bp);
        unsigned int q = bp;
        if (PSI) {
            // The reference compares psi of prefix + candidate (order.rs:84-99), first position on equal psi.  Does any
            // other pool point of the cluster lie in the band of running values where psi can coincide with the winner's?
            // (Every warp of every CTA folds the same table, so all of them take the branch together.)
            k2 = __reduce_min_sync(0xffffffffu, min(k2, (own_fp != kNone && own_fp != q) ? (unsigned int)(own_fk >> 32) : kNoKeyHi));
            const double a_best = __longlong_as_double((long long)(MINI ? bk : ~bk));
            const double band = order_tie_band(s_prefix, a_best, d);
            const double lim = MINI ? __dadd_rn(a_best, band) : __dsub_rn(a_best, band);
            if (k2 <= order_key_hi<MINI>(lim)) {
                __syncthreads();  // the running values written by the other warps of this CTA
                const double n_pairs = order_pairs(m + 1), inv_d = 1.0 / (double)d;
                unsigned long long pk = ~0ull;
                unsigned int pp = kNone;
                for (unsigned int s = tid; s < live; s += T) {
                    const double a = acc[s];
                    if (MINI ? a <= lim : a >= lim) {
                        const unsigned long long k = value_key<MINI>(order_psi(s_prefix, a, n_pairs, inv_d));
                        const unsigned int pos = s * kCl + rank;
                        if (k < pk || (k == pk && pos < pp)) { pk = k; pp = pos; }
                    }
                }
                warp_argmin_key(pk, pp);
                if (lane < kCl) {
                    CandRec* dst = cluster.map_shared_rank(cand_psi, lane) + rank * NW + warp;
                    dst->v = __longlong_as_double((long long)pk);
                    dst->p = (unsigned long long)pp;
                }
                cluster.sync();
                pk = ~0ull; pp = kNone;
#pragma unroll
                for (unsigned int j = 0; j < PER_LANE; ++j) {
                    const CandRec rec = cand_psi[(lane + 32 * j) % NREC];
                    const unsigned long long k = (unsigned long long)__double_as_longlong(rec.v);
                    const unsigned int pos = (unsigned int)rec.p;
                    if (k < pk || (k == pk && pos < pp)) { pk = k; pp = pos; }
                }
                warp_argmin_key(pk, pp);
                q = pp;
            }
        }
        // this table's phase is consumed by this warp: arm the phase of step + 2.  The warps still waiting look
        // for the parity of the completed phase, which cannot change before every warp of the cluster has sent
        // the records of step + 2.
        if (tid == 0) mbar_arrive_expect_tx(cbar, NREC * (unsigned int)sizeof(CandRec));
        OAS_T(3)
        const unsigned int L = pool_n - 1;
        const unsigned int q_owner = q % kCl, q_slot = q / kCl;
        const unsigned int l_owner = L % kCl, l_slot = L / kCl;
        if (tid <= d) {
            const double* src = tid < d ? coords + (size_t)q_slot * ds + tid : acc + q_slot;
            last_s[tid] = __longlong_as_double((long long)ld_cluster_b64(map_rank(smem_u32(src), q_owner)));
        }
        // ---- 3. the last entry moves into the hole: mail it to q's owner (from the far end of the CTA, so
        //         the warps that pull are not the ones that mail)
        if (L != q && rank == l_owner && tid >= T - (d + 2)) {
            const unsigned int e = tid - (T - (d + 2));
            const unsigned long long w = e < d    ? (unsigned long long)__double_as_longlong(coords[(size_t)l_slot * ds + e])
                                         : e == d ? (unsigned long long)__double_as_longlong(acc[l_slot])
                                                  : (unsigned long long)ids[l_slot];
            st_async_b64(map_rank(smem_u32(&mail[e]), q_owner), w, map_rank(hole_bar, q_owner));
        }
        OAS_T(4)
        __syncthreads();  // last_s complete: the pull of this CTA has landed
        OAS_T(5)
#if OAS_PROFILE
        if (tid == 0) t_ref_s = clock64();
#endif
        if (tid == 31) {
            if (rank == q_owner) p.perm[m] = ids[q_slot];
            if (L != q && rank != q_owner) mbar_arrive_remote(map_rank(hole_bar, q_owner));  // the owner's own arrival is armed
        }
        hole_slot = (L != q && rank == q_owner) ? q_slot : kNone;
        if (!MAXPRO && !CENTRE) cur_min = fmin(cur_min, last_s[d]);
        if (PSI) s_prefix = __dadd_rn(s_prefix, last_s[d]);
        --pool_n; ++m; ++step;
        OAS_T(6)
    };

    greedy_step(std::true_type{});
    while (pool_n > 0) greedy_step(std::false_type{});
#if OAS_PROFILE
    if (lane == 0 && rank == 5)
        printf("warp %u steps %u: slots done %lld, record sent %lld, table complete %lld, folded %lld, pull/mail issued %lld, (barrier passed: next step) , tail %lld\n",
               warp, step, tsum[0] / step, tsum[1] / step, tsum[2] / step, tsum[3] / step, tsum[4] / step, tsum[6] / step);
#endif
    cluster.sync();  // nobody leaves while its shared memory can still be addressed
}

size_t order_async_smem(unsigned long long n, unsigned long long d, int thr, int cl) {
    const unsigned long long slots = (n + cl - 1) / cl;
    const unsigned long long nrec = (unsigned long long)cl * (thr / 32);
    return (size_t)(3 * nrec * sizeof(CandRec) + 4 * 8 + slots * (d | 1ull) * 8 + slots * 8 + (d + 1) * 8 + (d + 2) * 8 + slots * 4 + 16);
}

constexpr int kOasThreads = 256;  // 7 worker warps + the hole warp
constexpr int kOasSlotsPerPass = 2;
// Measured against 384 / 512 threads and three / four slots per pass on (5000,8), (500,5), (1000,20), (10000,2):
// all within 5 %, this one ahead on three of the four shapes (profiles/r1j_order_ab.txt).

}  // namespace

bool order_cluster_async_supported(unsigned long long n, unsigned long long d) {
    return n >= 64 && n < (1ull << 31) && d >= 1 && d + 2 <= 224 && order_async_smem(n, d, kOasThreads, 8) + 1024 <= kSmemBudget;
}

template <int CL>
static cudaError_t launch_cluster(const OrderLaunch& L) {
    OrdAsParams p;
    p.x = L.design; p.n = (unsigned int)L.n; p.d = (unsigned int)L.d; p.ds = (unsigned int)(L.d | 1ull);
    p.slots = (unsigned int)((L.n + CL - 1) / CL);
    p.perm = L.perm;
    const size_t smem = order_async_smem(L.n, L.d, kOasThreads, CL);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (CL > 8) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL); cfg.blockDim = dim3(kOasThreads); cfg.dynamicSmemBytes = smem; cfg.stream = L.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, p);  // one cluster
        if (e == cudaSuccess) count_launch();
        return e;
    };
    if (L.metric == 0)
        return L.minimize ? go(order_cluster_async_kernel<true, true, kOasThreads, kOasSlotsPerPass, CL>)
                          : go(order_cluster_async_kernel<true, false, kOasThreads, kOasSlotsPerPass, CL>);
    return L.minimize ? go(order_cluster_async_kernel<false, true, kOasThreads, kOasSlotsPerPass, CL>)
                      : go(order_cluster_async_kernel<false, false, kOasThreads, kOasSlotsPerPass, CL>);
}

cudaError_t launch_order_cluster_async(const OrderLaunch& L) {
    // 16 CTAs halve the slots of a worker; worth the doubled record table (128 records folded by every warp)
    // once a worker of an 8-CTA cluster would hold three or more slots.
    int cl = (L.n + 7) / 8 > 2ull * (kOasThreads - 32) ? 16 : 8;
    if (const char* v = cfg("MAXPRO_ORDER_CLUSTER_CTAS")) cl = atoi(v) == 16 ? 16 : 8;
    if (cl == 16) {
        cudaError_t e = launch_cluster<16>(L);
        if (e == cudaSuccess) return e;
        (void)cudaGetLastError();  // a 16-CTA cluster could not be placed: the portable size always can
    }
    return launch_cluster<8>(L);
}

}  // namespace mp

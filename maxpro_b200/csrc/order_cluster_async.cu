// K4 on a thread-block cluster WITHOUT cluster barriers in the step: the exchanges of
// order_cluster.cu done with mbarriers, st.async (remote store that completes bytes on the
// destination's mbarrier) and remote loads through distributed shared memory.
//
// Same algorithm, pool layout and tie rules as order_cluster.cu (src/order.rs:34-105 with one
// running value per remaining point; scan position p lives in CTA p mod 8, slot p / 8).  That
// kernel pays two `cluster.sync` per greedy step (~380 cycles each, every thread of 8 CTAs) plus a
// remote store in front of each.  Here a step is
//
//   1  every CTA updates its slots against the point just placed and finds its best (value, position);
//      eight lanes send that 16-byte record to the candidate table of every CTA with st.async, which
//      counts the bytes on the DESTINATION's mbarrier; each CTA waits on its own mbarrier only
//      (tables and mbarriers double-buffered by step parity);
//   2  every CTA folds the 8 records (same answer everywhere: position q wins) and PULLS the winner's
//      coordinates (and running value) out of its owner's shared memory with d + 1 remote loads;
//   3  off the critical path: the owner of the last position L mails its entry to the owner of q
//      (Vec::swap_remove moves the last element into the hole, order.rs:63,101) with st.async on the
//      owner's `hole` mbarrier, and every CTA arrives on that mbarrier once its pull has landed.  The
//      owner fills the hole when the mbarrier completes (mail in, nobody still reading the slot) --
//      one thread does it after its regular slots of the NEXT step, so the wait is normally over.
//
// Critical path per step: update loop, block argbest, one DSMEM store latency, fold, one DSMEM load
// latency.  No thread ever waits for a thread of another CTA except through those two.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace mp {

namespace {

constexpr int kCl = 8;          // CTAs per cluster (portable maximum)
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

// plain remote store / volatile local load of one 16-byte record (flag-polling exchange)
__device__ __forceinline__ void st_cluster_v2(unsigned int remote_addr, unsigned long long a, unsigned long long b) {
    asm volatile("st.shared::cluster.v2.b64 [%0], {%1, %2};" ::"r"(remote_addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2(unsigned int addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}

#ifndef OAS_PROFILE
#define OAS_PROFILE 0
#endif
#if OAS_PROFILE
#define OAS_T(i) { const long long now_ = clock64(); tsum[i] += now_ - tprev; tprev = now_; }
#else
#define OAS_T(i)
#endif

// XCH: how the per-step records travel.  0 = st.async + mbarrier (transaction bytes), one record per CTA;
//      1 = plain remote stores of self-validating words, receivers poll their own shared memory, one record per CTA;
//      2 = as 1 but one record per WARP, no block-level reduction and no CTA barrier before the exchange.
// LOOPV: 0 = dimension loops unrolled by the compiler, two slots evaluated per pass (a lone slot twice);
//        1 = dimension loops not unrolled (small code), a lone slot evaluated once.
template <bool MAXPRO, bool MINIMIZE, int THR, int XCH, int LOOPV>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(THR, 1) order_cluster_async_kernel(OrdAsParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank();
    const unsigned int n = p.n, d = p.d, ds = p.ds, slots = p.slots;
    const unsigned int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned int T = THR, NW = THR / 32, TW = THR - 32;  // TW worker threads; the last warp serves the hole
    constexpr unsigned int RPC = XCH == 2 ? NW : 1;         // records per CTA and step
    constexpr unsigned int NREC = kCl * RPC;                // records per table
    constexpr unsigned int PER_LANE = (NREC + 31) / 32;
    static_assert(NW <= 32, "second reduction level is one warp");

    extern __shared__ __align__(16) unsigned char oas_smem[];
    CandRec* cand = reinterpret_cast<CandRec*>(oas_smem);                     // [2][NREC] candidate tables by step parity
    CandRec* red = cand + 2 * NREC;                                           // [NW]     per-warp bests
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + NW);  // [4]  cand mbarriers (2), hole mbarrier, pad
    double* coords = reinterpret_cast<double*>(bars + 4);                     // [slots][ds]
    double* acc = coords + (size_t)slots * ds;                                // [slots]
    double* last_s = acc + slots;                                             // [d + 1]  winner: coordinates + running value
    unsigned long long* mail = reinterpret_cast<unsigned long long*>(last_s + d + 1);  // [d + 2]  coordinates, running value, id
    unsigned int* ids = reinterpret_cast<unsigned int*>(mail + d + 2);        // [slots]

    const unsigned int cand_bar0 = smem_u32(&bars[0]), hole_bar = smem_u32(&bars[2]);

    if (tid == 0) {
        mbar_init(cand_bar0, 1);       // thread 0's arrive.expect_tx; the 8 records arrive as transaction bytes
        mbar_init(cand_bar0 + 8, 1);
        mbar_init(hole_bar, kCl);      // one arrival per CTA (pull finished) + the mailed entry as transaction bytes
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (unsigned int e = tid; e < 2 * NREC; e += T) { cand[e].v = 0.0; cand[e].p = 0ull; }  // tag 0 = nothing yet
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
#if OAS_PROFILE
    long long tsum[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    // One greedy step.  CENTRE = the first point: distance to the centre of the cube, smaller wins, first
    // index on ties (order.rs:49-59); afterwards: the running value against the point just placed (last_s).
    auto greedy_step = [&](auto centre_tag) {
        constexpr bool CENTRE = decltype(centre_tag)::value;
        constexpr bool MINI = CENTRE ? true : MINIMIZE;
        constexpr int UNR = LOOPV ? 1 : 4;
#if OAS_PROFILE
        long long tprev = clock64();
#endif
        // values of two slots at once (independent dependency chains)
        auto eval2 = [&](unsigned int s0, unsigned int s1, double& v0, double& v1) {
            const double* x0 = coords + (size_t)s0 * ds;
            const double* x1 = coords + (size_t)s1 * ds;
            if (CENTRE) {
                double f0 = __dsub_rn(0.5, x0[0]), f1 = __dsub_rn(0.5, x1[0]);
                double q0 = __dmul_rn(f0, f0), q1 = __dmul_rn(f1, f1);
#pragma unroll UNR
                for (unsigned int k = 1; k < d; ++k) {
                    f0 = __dsub_rn(0.5, x0[k]); f1 = __dsub_rn(0.5, x1[k]);
                    q0 = __dadd_rn(q0, __dmul_rn(f0, f0)); q1 = __dadd_rn(q1, __dmul_rn(f1, f1));
                }
                v0 = sqrt(q0); v1 = sqrt(q1);
            } else if (MAXPRO) {  // src/maxpro_utils.rs:45-54, one pair
                double p0 = 1.0, p1 = 1.0;
#pragma unroll UNR
                for (unsigned int k = 0; k < d; ++k) {
                    const double l = last_s[k];
                    const double f0 = __dsub_rn(l, x0[k]), f1 = __dsub_rn(l, x1[k]);
                    p0 = __dmul_rn(p0, __dmul_rn(f0, f0)); p1 = __dmul_rn(p1, __dmul_rn(f1, f1));
                }
                v0 = __dadd_rn(acc[s0], 1.0 / __dadd_rn(p0, kEps));
                v1 = __dadd_rn(acc[s1], 1.0 / __dadd_rn(p1, kEps));
            } else {
                double l = last_s[0];
                double f0 = __dsub_rn(l, x0[0]), f1 = __dsub_rn(l, x1[0]);
                double q0 = __dmul_rn(f0, f0), q1 = __dmul_rn(f1, f1);
#pragma unroll UNR
                for (unsigned int k = 1; k < d; ++k) {
                    l = last_s[k];
                    f0 = __dsub_rn(l, x0[k]); f1 = __dsub_rn(l, x1[k]);
                    q0 = __dadd_rn(q0, __dmul_rn(f0, f0)); q1 = __dadd_rn(q1, __dmul_rn(f1, f1));
                }
                v0 = fmin(acc[s0], sqrt(q0)); v1 = fmin(acc[s1], sqrt(q1));
            }
        };
        // one slot; coordinates and old running value passed in (the hole is evaluated out of the mailbox)
        auto eval1 = [&](const double* x0, double a0) -> double {
            if (CENTRE) {
                double f0 = __dsub_rn(0.5, x0[0]);
                double q0 = __dmul_rn(f0, f0);
#pragma unroll UNR
                for (unsigned int k = 1; k < d; ++k) {
                    f0 = __dsub_rn(0.5, x0[k]);
                    q0 = __dadd_rn(q0, __dmul_rn(f0, f0));
                }
                return sqrt(q0);
            } else if (MAXPRO) {
                double p0 = 1.0;
#pragma unroll UNR
                for (unsigned int k = 0; k < d; ++k) {
                    const double f0 = __dsub_rn(last_s[k], x0[k]);
                    p0 = __dmul_rn(p0, __dmul_rn(f0, f0));
                }
                return __dadd_rn(a0, 1.0 / __dadd_rn(p0, kEps));
            } else {
                double f0 = __dsub_rn(last_s[0], x0[0]);
                double q0 = __dmul_rn(f0, f0);
#pragma unroll UNR
                for (unsigned int k = 1; k < d; ++k) {
                    f0 = __dsub_rn(last_s[k], x0[k]);
                    q0 = __dadd_rn(q0, __dmul_rn(f0, f0));
                }
                return fmin(a0, sqrt(q0));
            }
        };
        unsigned long long bk = ~0ull;
        unsigned int bp = kNone;
        // keep the running value, offer (value, position) to the argbest
        auto offer = [&](unsigned int s, double v) {
            double c = v;
            if (!CENTRE) {
                acc[s] = v;
                if (!MAXPRO) c = fmin(v, cur_min);  // maximin: criterion of prefix + this point
            }
            const unsigned long long k = value_key<MINI>(c);
            const unsigned int pos = s * kCl + rank;
            if (k < bk || (k == bk && pos < bp)) { bk = k; bp = pos; }
        };

        const unsigned int par = step & 1u;
        const unsigned int cbar = cand_bar0 + 8 * par;
        // the arrival this step's records are counted against: off the critical path (its previous phase, two
        // steps ago, is complete -- this thread waited for it)
        if (XCH != 1 && tid == 0) mbar_arrive_expect_tx(cbar, NREC * (unsigned int)sizeof(CandRec));
        // ---- 1a. the worker warps update this CTA's slots
        const unsigned int live = pool_n > rank ? (pool_n - rank + kCl - 1) / kCl : 0;  // slots with position < pool_n
        if (tid < TW) {
            for (unsigned int s0 = tid; s0 < live; s0 += 2 * TW) {
                const unsigned int s1 = s0 + TW;
                const bool ok0 = s0 != hole_slot, ok1 = s1 < live && s1 != hole_slot;
                if (LOOPV == 0 || (ok0 && ok1)) {
                    double v0, v1;
                    eval2(s0, ok1 ? s1 : s0, v0, v1);
                    if (ok0) offer(s0, v0);
                    if (ok1) offer(s1, v1);
                } else if (ok0) {
                    offer(s0, eval1(coords + (size_t)s0 * ds, acc[s0]));
                } else if (ok1) {
                    offer(s1, eval1(coords + (size_t)s1 * ds, acc[s1]));
                }
            }
        }
        OAS_T(0)
        // ---- 1b. the last warp has no slots of its own: it fills the hole left by the previous winner once the
        //          mail is in and every CTA has pulled the winner's coordinates out of the slot.  Lane 0 evaluates
        //          the entry straight out of the mailbox while the other lanes copy it into the slot.
        if (hole_slot != kNone) {
            if (tid >= TW) {
                mbar_wait(hole_bar, hole_phase);
                double* xc = coords + (size_t)hole_slot * ds;
                for (unsigned int k = lane; k < d; k += 32) xc[k] = __longlong_as_double((long long)mail[k]);
                if (lane == 1 % 32) ids[hole_slot] = (unsigned int)mail[d + 1];
                if (lane == 0) offer(hole_slot, eval1(reinterpret_cast<const double*>(mail), __longlong_as_double((long long)mail[d])));
                __syncwarp();
            }
            hole_phase ^= 1u;
        }
        OAS_T(1)
        // ---- 1c. argbest of the warp, (XCH < 2: of the block,) records to every CTA of the cluster
        const unsigned long long tag = (step + 1u) & 0xFFFFu;
        warp_argmin_key(bk, bp);
        // self-validating words: [key half 32][position half 16][tag 16] -- each 8-byte word can be checked alone
        auto word0 = [&](unsigned long long k, unsigned int pos) { return (k & 0xFFFFFFFF00000000ull) | ((unsigned long long)(pos >> 16) << 16) | tag; };
        auto word1 = [&](unsigned long long k, unsigned int pos) { return (k << 32) | ((unsigned long long)(pos & 0xFFFFu) << 16) | tag; };
        if (XCH == 2) {
            if (lane < kCl)  // lane r stores this warp's record into CTA r
                st_async_v2(map_rank(smem_u32(&cand[par * NREC + rank * RPC + warp]), lane), bk, (unsigned long long)bp, map_rank(cbar, lane));
            OAS_T(2) OAS_T(3) OAS_T(4)
        } else {
            if (lane == 0) { red[warp].v = __longlong_as_double((long long)bk); red[warp].p = bp; }
            OAS_T(2)
            __syncthreads();
            OAS_T(3)
            if (warp == 0) {
                bk = lane < NW ? (unsigned long long)__double_as_longlong(red[lane].v) : ~0ull;
                bp = lane < NW ? (unsigned int)red[lane].p : kNone;
                warp_argmin_key(bk, bp);
                if (XCH == 0) {
                    if (lane < kCl)  // lane r stores into CTA r
                        st_async_v2(map_rank(smem_u32(&cand[par * NREC + rank]), lane), bk, (unsigned long long)bp, map_rank(cbar, lane));
                } else if (lane < kCl) {
                    st_cluster_v2(map_rank(smem_u32(&cand[par * NREC + rank]), lane), word0(bk, bp), word1(bk, bp));
                }
            }
            OAS_T(4)
        }
        // ---- 2. wait for the records, fold them (same answer in every warp of every CTA), pull the winner
        if (XCH != 1) {
            mbar_wait(cbar, (step >> 1) & 1u);
            OAS_T(5)
            bk = ~0ull; bp = kNone;
#pragma unroll
            for (unsigned int j = 0; j < PER_LANE; ++j) {
                const CandRec rec = cand[par * NREC + (lane + 32 * j) % NREC];
                const unsigned long long k = (unsigned long long)__double_as_longlong(rec.v);
                const unsigned int pos = (unsigned int)rec.p;
                if (k < bk || (k == bk && pos < bp)) { bk = k; bp = pos; }
            }
        } else {
            unsigned long long w0[PER_LANE], w1[PER_LANE];
            unsigned int missing = (1u << PER_LANE) - 1u;
            unsigned long long t0 = 0;
            for (unsigned int spin = 1;; ++spin) {
#pragma unroll
                for (unsigned int j = 0; j < PER_LANE; ++j) {
                    if (missing & (1u << j)) {
                        const unsigned int idx = (lane + 32 * j) % NREC;
                        ld_volatile_v2(smem_u32(&cand[par * NREC + idx]), w0[j], w1[j]);
                        if ((w0[j] & 0xFFFFu) == tag && (w1[j] & 0xFFFFu) == tag) missing &= ~(1u << j);
                    }
                }
                if (__all_sync(0xffffffffu, missing == 0)) break;
                if ((spin & 0x3FFu) == 0) {  // a protocol error must not hang the device
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > 2000000000ull) __trap();
                }
            }
            OAS_T(5)
            bk = ~0ull; bp = kNone;
#pragma unroll
            for (unsigned int j = 0; j < PER_LANE; ++j) {
                const unsigned long long k = (w0[j] & 0xFFFFFFFF00000000ull) | (w1[j] >> 32);
                const unsigned int pos = (unsigned int)(((w0[j] >> 16) & 0xFFFFu) << 16) | (unsigned int)((w1[j] >> 16) & 0xFFFFu);
                if (k < bk || (k == bk && pos < bp)) { bk = k; bp = pos; }
            }
        }
        warp_argmin_key(bk, bp);
        const unsigned int q = bp;
        OAS_T(6)
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
        OAS_T(7)
        __syncthreads();  // last_s complete: the pull of this CTA has landed
        OAS_T(8)
        if (tid == T - 1) {  // not warp 0: it sends the records of the next step
            if (rank == q_owner) p.perm[m] = ids[q_slot];
            if (L != q) {
                if (rank == q_owner) mbar_arrive_expect_tx(hole_bar, (d + 2) * 8u);
                else mbar_arrive_remote(map_rank(hole_bar, q_owner));
            }
        }
        hole_slot = (L != q && rank == q_owner) ? q_slot : kNone;
        if (!MAXPRO && !CENTRE) cur_min = fmin(cur_min, last_s[d]);
        --pool_n; ++m; ++step;
        OAS_T(9)
    };

    greedy_step(std::true_type{});
    while (pool_n > 0) greedy_step(std::false_type{});
#if OAS_PROFILE
    if ((tid == 0 || tid == T - 1) && (rank == 0 || rank == 5))
        printf("rank %u tid %u steps %u: loop %lld hole %lld argmin %lld sync %lld publish %lld wait %lld fold %lld pull+mail %lld sync %lld tail %lld\n", rank,
               tid, step, tsum[0] / step, tsum[1] / step, tsum[2] / step, tsum[3] / step, tsum[4] / step, tsum[5] / step, tsum[6] / step,
               tsum[7] / step, tsum[8] / step, tsum[9] / step);
#endif
    cluster.sync();  // nobody leaves while its shared memory can still be addressed
}

size_t order_async_smem(unsigned long long n, unsigned long long d, int thr, int xch) {
    const unsigned long long slots = (n + kCl - 1) / kCl;
    const unsigned long long nw = thr / 32, nrec = kCl * (xch == 2 ? nw : 1);
    return (size_t)(2 * nrec * sizeof(CandRec) + nw * sizeof(CandRec) + 4 * 8 + slots * (d | 1ull) * 8 + slots * 8 +
                    (d + 1) * 8 + (d + 2) * 8 + slots * 4 + 16);
}

}  // namespace

bool order_cluster_async_supported(unsigned long long n, unsigned long long d) {
    return n >= 64 && n < (1ull << 31) && d >= 1 && d + 2 <= 224 && order_async_smem(n, d, 512, 2) + 1024 <= kSmemBudget;
}

template <int THR, int XCH, int LOOPV>
static cudaError_t launch_variant(const OrderLaunch& L) {
    OrdAsParams p;
    p.x = L.design; p.n = (unsigned int)L.n; p.d = (unsigned int)L.d; p.ds = (unsigned int)(L.d | 1ull);
    p.slots = (unsigned int)((L.n + kCl - 1) / kCl);
    p.perm = L.perm;
    const size_t smem = order_async_smem(L.n, L.d, THR, XCH);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<kCl, THR, smem, L.stream>>>(p);  // one cluster (compile-time __cluster_dims__)
        count_launch();
        return cudaGetLastError();
    };
    if (L.metric == 0)
        return L.minimize ? go(order_cluster_async_kernel<true, true, THR, XCH, LOOPV>) : go(order_cluster_async_kernel<true, false, THR, XCH, LOOPV>);
    return L.minimize ? go(order_cluster_async_kernel<false, true, THR, XCH, LOOPV>) : go(order_cluster_async_kernel<false, false, THR, XCH, LOOPV>);
}

cudaError_t launch_order_cluster_async(const OrderLaunch& L) {
    int thr = 256, xch = 2, loopv = 0;
    if (const char* v = getenv("MAXPRO_OAS_VARIANT")) sscanf(v, "%d,%d,%d", &thr, &xch, &loopv);
#define OAS_CASE(T_, X_, V_) if (thr == T_ && xch == X_ && loopv == V_) return launch_variant<T_, X_, V_>(L);
    OAS_CASE(256, 0, 0) OAS_CASE(256, 0, 1) OAS_CASE(256, 1, 0) OAS_CASE(256, 1, 1) OAS_CASE(256, 2, 0) OAS_CASE(256, 2, 1)
    OAS_CASE(512, 0, 0) OAS_CASE(512, 0, 1) OAS_CASE(512, 1, 0) OAS_CASE(512, 1, 1) OAS_CASE(512, 2, 0) OAS_CASE(512, 2, 1)
#undef OAS_CASE
    return cudaErrorInvalidValue;
}

}  // namespace mp

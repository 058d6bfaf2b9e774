// K3, swap moves on the maximin criterion, scored incrementally and exactly.
//
// Replaces anneal::anneal_lhd with swap = true and maximin_criterion as the metric
// (src/anneal.rs:56-143, src/maximin_utils.rs:54-67).  The reference re-scores all n(n-1)/2
// pairs after every move; anneal.cu does the same on the GPU (O(n^2 d) per step).  Here the
// chain carries, for every row i, the squared distance to its nearest neighbour and who that
// neighbour is:
//     rowmin[i] = min_{j != i} |x_i - x_j|^2 ,   rowarg[i] = a j that attains it.
// The criterion is sqrt(min_i rowmin[i]): the reference takes sqrt per pair and then the minimum,
// and sqrt is monotone and correctly rounded, so the two agree bit for bit.  Exchanging x[r1][c]
// and x[r2][c] changes only the pairs that touch r1 or r2 (the pair (r1, r2) itself keeps its
// distance: one difference changes sign).  For a row i outside {r1, r2}:
//   * rowarg[i] not in {r1, r2}: the old minimum is still attained, so
//       rowmin'[i] = min(rowmin[i], |x_i - x'_r1|^2, |x_i - x'_r2|^2);
//   * otherwise the row is "dirty": its nearest neighbour moved, and the minimum over all j is
//     taken again (about two rows per move: a point is the nearest neighbour of one other on average).
// rowmin'[r1], rowmin'[r2] are the minima of the new distances every row has just computed.
// Every distance is evaluated in full, in the reference's order (maximin_utils.rs:39-44: separate
// multiply and add, dimensions left to right), never updated by subtraction, and a minimum does
// not depend on the order it is taken in -- so the chain follows the reference's trajectory
// exactly at O(n d) instead of O(n^2 d) per step.
// The tentative rowmin'/rowarg' go into the second of two buffers; an accepted move flips them.
// Draws, Metropolis rule, global best and the move log are those of anneal.cu.
#include "anneal.cuh"

#include <algorithm>
#include <cstdlib>

#ifndef AMX_PROFILE
#define AMX_PROFILE 0  // 1: clock64 sums per phase of the step, printed by thread 0 of CTA 0 (make EXTRA=-DAMX_PROFILE=1)
#endif
#if AMX_PROFILE
#include <cstdio>
__device__ long long g_amx_trace[4 * 8 * 16];  // [step 500..503][probe][warp] clock of lane 0, CTA 0, chain 0 (screened kernel)
#define AMX_T(i) { const long long now_ = clock64(); tsum[i] += now_ - tprev; tprev = now_; \
                   if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && amx_t >= 500 && amx_t < 504) g_amx_trace[((amx_t - 500) * 8 + (i)) * 16 + (threadIdx.x >> 5)] = now_; }
#else
#define AMX_T(i)
#endif

namespace mp {

constexpr int kMaxDirty = 6;  // dirty rows re-scored per pass

struct MinArg {
    double v;
    int a;
    // strict "<" on the bit patterns: squared distances are >= +0, so the integer order is the numeric order, and
    // the compare runs on the integer pipes instead of a DSETP on the FP64 pipe (8 of these per row of the pass)
    __device__ __forceinline__ void take(double w, int b) {
        if ((unsigned long long)__double_as_longlong(w) < (unsigned long long)__double_as_longlong(v)) { v = w; a = b; }
    }
};

// Warp minimum of (squared distance, row), smaller row on equal distances.  A squared distance is >= +0, so its
// bit pattern orders like the number: three REDUX.MIN (high word, low word among the lanes that hold the high
// minimum, row among the lanes that hold both) instead of five shuffle levels of two-DSETP compares -- the
// shuffle tree of such a pair costs ~600 cycles of latency per call (measured in order_cluster_async.cu) and
// its compares run on the FP64 pipe, which is what this kernel is short of.
__device__ __forceinline__ MinArg warp_minarg(MinArg m) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(m.v);
    const unsigned int hi = (unsigned int)(bits >> 32), lo = (unsigned int)bits;
    const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    const bool mine = hi == mhi && lo == mlo;
    const unsigned int ma = __reduce_min_sync(0xffffffffu, mine ? (unsigned int)m.a : 0xFFFFFFFFu);
    m.v = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
    m.a = (int)ma;
    return m;
}

// |a - x_j|^2 in the reference's order; a[] is a broadcast row (stride `as` doubles), x_j is column j of xs
__device__ __forceinline__ double sqdist_row(const double* a, int as, const double* xs, int ns, int j, int d) {
    double df = __dsub_rn(a[0], xs[j]);
    double s = __dmul_rn(df, df);
    for (int k = 1; k < d; ++k) {
        df = __dsub_rn(a[k * as], xs[k * ns + j]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

// Squared distances from row j to the two moved rows (ab[k] = {x'_r1[k], x'_r2[k]}, a 16-byte broadcast)
// and to CNT dirty rows (columns di[q] of xs, broadcast reads), in one walk over the dimensions:
// x_j[k] is read once for all of them.  Each sum is formed in the reference's order.
template <int CNT>
__device__ __forceinline__ void sqdist_fused(const double* ab, const double* xs, int ns, int j, int d, const int (&di)[kMaxDirty],
                                             double& sa, double& sb, double (&sq)[kMaxDirty]) {
    const double* xj = xs + j;
    const double* xq[CNT > 0 ? CNT : 1];
#pragma unroll
    for (int q = 0; q < CNT; ++q) xq[q] = xs + di[q];
    {
        const double2 r = *reinterpret_cast<const double2*>(ab);
        const double x = *xj;
        const double da = __dsub_rn(r.x, x), db = __dsub_rn(r.y, x);
        sa = __dmul_rn(da, da); sb = __dmul_rn(db, db);
#pragma unroll
        for (int q = 0; q < CNT; ++q) { const double dq = __dsub_rn(*xq[q], x); sq[q] = __dmul_rn(dq, dq); }
    }
#pragma unroll 2
    for (int k = 1; k < d; ++k) {
        xj += ns; ab += 2;
        const double2 r = *reinterpret_cast<const double2*>(ab);
        const double x = *xj;
        const double da = __dsub_rn(r.x, x), db = __dsub_rn(r.y, x);
        sa = __dadd_rn(sa, __dmul_rn(da, da));
        sb = __dadd_rn(sb, __dmul_rn(db, db));
#pragma unroll
        for (int q = 0; q < CNT; ++q) { xq[q] += ns; const double dq = __dsub_rn(*xq[q], x); sq[q] = __dadd_rn(sq[q], __dmul_rn(dq, dq)); }
    }
}

// One pass of the block over the rows j outside {r1, r2}: new distances to the moved rows (first
// chunk: row minima of the clean rows and the running minima of rows r1, r2) and to the CNT dirty
// rows of this chunk.
template <int CNT>
__device__ __forceinline__ void maximin_row_pass(const double* cur, const double* rowsx, int n, int d, int ns, int r1, int r2, bool first,
                                                 const double* rm, const unsigned short* ra, double* rm_new, unsigned short* ra_new,
                                                 const int (&di)[kMaxDirty], MinArg& all, MinArg& m1, MinArg& m2, MinArg (&dm)[kMaxDirty]) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        if (j == r1 || j == r2) continue;
        double dn1, dn2, sq[kMaxDirty];
        sqdist_fused<CNT>(rowsx, cur, ns, j, d, di, dn1, dn2, sq);
        if (first) {
            m1.take(dn1, j);
            m2.take(dn2, j);
            const int a = ra[j];
            if (a != r1 && a != r2) {  // clean row: the old minimum is still attained
                MinArg mj{rm[j], a};
                mj.take(dn1, r1);
                mj.take(dn2, r2);
                rm_new[j] = mj.v; ra_new[j] = (unsigned short)mj.a;
                all.take(mj.v, j);
            }
        }
#pragma unroll
        for (int q = 0; q < CNT; ++q) {
            if (j == di[q]) { dm[q].take(dn1, r1); dm[q].take(dn2, r2); }  // its distances to the moved rows
            else dm[q].take(sq[q], j);
        }
    }
}

// minimum of the per-warp slots of one block minimum (bit patterns of non-negative doubles): a load per lane, two REDUX.MIN
__device__ __forceinline__ unsigned long long block_min_bits(const unsigned long long* slots, int nwarps, int lane) {
    const unsigned long long m = lane < nwarps ? slots[lane] : 0x7FF0000000000000ull;
    const unsigned int hi = (unsigned int)(m >> 32), lo = (unsigned int)m;
    const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    return ((unsigned long long)mhi << 32) | mlo;
}

__global__ void __launch_bounds__(512) anneal_maximin_swap_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;
    const int dns = d * ns;
    double* cur = smem_d;                                     // [k][i]
    double* rmin0 = cur + dns;                                // [2][ns]: committed / tentative row minima
    double* rowsx = rmin0 + 2 * ns;                           // [d][2]: the two moved rows AFTER the move
    double* sqk = rowsx + 2 * d;                              // [d] squared differences of the pair (r1, r2), then [d] = their sum
    // per-warp minima (bit patterns), [slot][16 warps]: slot 0 clean rows, 1 row r1, 2 row r2, 3 + q dirty row q of the chunk
    unsigned long long* lmin = reinterpret_cast<unsigned long long*>(sqk + d + 1);  // [3 + kMaxDirty][16]
    Decision* dec = reinterpret_cast<Decision*>(lmin + (3 + kMaxDirty) * 16);
    int* dirty_n = reinterpret_cast<int*>(dec + 1);           // dirty_n[0] = count, then the list
    int* dirty = dirty_n + 1;                                 // [n] worst case
    // aligned on the OFFSET: a pointer that went through an integer loses its address space, and everything behind it
    // would be read with generic loads (and generic atomics)
    const size_t draws_off = ((size_t)(reinterpret_cast<unsigned char*>(dirty + n) - reinterpret_cast<unsigned char*>(smem_d)) + 15) & ~(size_t)15;
    uint4* draws = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(smem_d) + draws_off);
    ushort4* mlog = reinterpret_cast<ushort4*>(draws + kDrawBatch);
    unsigned short* rarg0 = reinterpret_cast<unsigned short*>(mlog + kLogCap);  // [2][ns] their arguments
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)blockDim.x >> 5;

    __shared__ CtaBest cta_best;  // per-CTA fold of the chains' global bests (anneal.cuh)
    if (threadIdx.x == 0) cta_best_reset(&cta_best, p);
    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            const int i = e / d, k = e - i * d;
            cur[k * ns + i] = p.design[e];  // anneal.rs:89
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) { rmin0[i] = p.rowmin0[i]; rarg0[i] = p.rowarg0[i]; }
        if (threadIdx.x == 0) dirty_n[0] = 0;
        int sel = 0;  // which buffer holds the committed row minima
        double raw_cur = 0.0, cur_metric = 0.0, global_best = 0.0;
        double temp = p.t0;                                                             // anneal.rs:81
        bool materialized = false, overflow = false;
        int log_n = 0;
        if (warp == 0) {
            raw_cur = *p.raw0;
            cur_metric = sqrt(raw_cur);                                                 // :90
            global_best = cur_metric;                                                   // :94
        }
        double* best_out = cta_best_slot(p, &cta_best, nd);
        __syncthreads();

#if AMX_PROFILE
        long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#endif
        for (unsigned long long t = 0; t < p.iters; ++t) {
#if AMX_PROFILE
            const long long amx_t = chain == 0 ? (long long)t : -1;
#endif
            const unsigned int slot = (unsigned int)t & (kDrawBatch - 1);
            if (slot == 0) {
                for (int q = threadIdx.x; q < kDrawBatch; q += blockDim.x) {
                    const unsigned long long tq = t + (unsigned long long)q;
                    const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
                    // decoded once for all threads and steps of the batch -- anneal.rs:29-31: r1, r2, c; the accept draw stays raw
                    draws[q] = make_uint4(mulhi_bound(w.x, (uint32_t)n), mulhi_bound(w.y, (uint32_t)n), mulhi_bound(w.z, (uint32_t)d), w.w);
                }
                __syncthreads();
            }
            const uint4 w = draws[slot];
            const int r1 = (int)w.x, r2 = (int)w.y, c = (int)w.z;
            const double* rm = rmin0 + sel * ns;
            const unsigned short* ra = rarg0 + sel * ns;
            double* rm_new = rmin0 + (sel ^ 1) * ns;
            unsigned short* ra_new = rarg0 + (sel ^ 1) * ns;
            const bool moved = r1 != r2;

            // Block minima on the bit patterns (squared distances are >= +0, so the integer order is the numeric
            // order): every warp leaves its minimum in its own slot, and after the barrier every warp folds the (at most 16)
            // slots of a minimum with two REDUX.MIN -- a 64-bit atomicMin on shared memory is a compare-and-swap loop
            // (ATOMS.CAST.SPIN), 3 + dirty of them per warp on the same few addresses.  Who attains a minimum is written
            // afterwards by every warp whose partial equals it -- the decision needs the values only.
            AMX_T(0)
            double fin_all = CUDART_INF, g1 = CUDART_INF, g2 = CUDART_INF, d12 = CUDART_INF;
            if (moved) {
                // ---- stage the two rows as they are AFTER the exchange, list the dirty rows
                for (int k = threadIdx.x; k < d; k += blockDim.x) {
                    const double a = cur[k * ns + r1], b = cur[k * ns + r2];
                    const double a2 = (k == c) ? b : a, b2 = (k == c) ? a : b;
                    rowsx[2 * k] = a2;       // x'[r1][k]
                    rowsx[2 * k + 1] = b2;   // x'[r2][k]
                    const double df = __dsub_rn(a2, b2);
                    sqk[k] = __dmul_rn(df, df);
                }
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    const int a = ra[i];
                    if (i != r1 && i != r2 && (a == r1 || a == r2)) dirty[atomicAdd(dirty_n, 1)] = i;  // dirty_n was zeroed before the last barrier
                }
                __syncthreads();
                AMX_T(1)
                const int nd_rows = dirty_n[0];
                if (warp == 0) {
                    // the pair (r1, r2) itself: its squares summed in the reference's order
                    double s12 = sqk[0];
#pragma unroll 4
                    for (int k = 1; k < d; ++k) s12 = __dadd_rn(s12, sqk[k]);
                    if (lane == 0) sqk[d] = s12;
                }

                // ---- pass over the rows: new distances to r1 and r2, and to the dirty rows
                MinArg all{CUDART_INF, 0}, m1{CUDART_INF, 0}, m2{CUDART_INF, 0};
                for (int base = 0; base == 0 || base < nd_rows; base += kMaxDirty) {
                    const int cnt = min(kMaxDirty, nd_rows - base);
                    if (base != 0) __syncthreads();  // more than kMaxDirty dirty rows: the previous chunk's minima have been read
                    MinArg dm[kMaxDirty];
#pragma unroll
                    for (int q = 0; q < kMaxDirty; ++q) dm[q] = MinArg{CUDART_INF, 0};
                    int di[kMaxDirty];
#pragma unroll
                    for (int q = 0; q < kMaxDirty; ++q) di[q] = q < cnt ? dirty[base + q] : 0;
                    const bool first = base == 0;
                    switch (cnt) {
                        case 0: maximin_row_pass<0>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        case 1: maximin_row_pass<1>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        case 2: maximin_row_pass<2>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        case 3: maximin_row_pass<3>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        case 4: maximin_row_pass<4>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        case 5: maximin_row_pass<5>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                        default: maximin_row_pass<6>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm); break;
                    }
                    AMX_T(2)
                    // ---- warp minima, then the block minima by atomicMin
                    if (first) {
                        all = warp_minarg(all); m1 = warp_minarg(m1); m2 = warp_minarg(m2);
                        if (lane == 0) {
                            lmin[0 * 16 + warp] = (unsigned long long)__double_as_longlong(all.v);
                            lmin[1 * 16 + warp] = (unsigned long long)__double_as_longlong(m1.v);
                            lmin[2 * 16 + warp] = (unsigned long long)__double_as_longlong(m2.v);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kMaxDirty; ++q) {
                        if (q < cnt) {
                            dm[q] = warp_minarg(dm[q]);
                            if (lane == 0) lmin[(3 + q) * 16 + warp] = (unsigned long long)__double_as_longlong(dm[q].v);
                        }
                    }
                    AMX_T(3)
                    __syncthreads();
                    AMX_T(4)
                    // the dirty rows of this chunk: value and a neighbour that attains it
#pragma unroll
                    for (int q = 0; q < kMaxDirty; ++q) {
                        if (q < cnt) {
                            const unsigned long long g = block_min_bits(lmin + (3 + q) * 16, nwarps, lane);
                            if (lane == 0 && (unsigned long long)__double_as_longlong(dm[q].v) == g) { rm_new[di[q]] = dm[q].v; ra_new[di[q]] = (unsigned short)dm[q].a; }
                            fin_all = fmin(fin_all, __longlong_as_double((long long)g));
                        }
                    }
                }
                // rows r1 and r2: the minimum over the other rows and the pair (r1, r2) itself
                fin_all = fmin(fin_all, __longlong_as_double((long long)block_min_bits(lmin + 0 * 16, nwarps, lane)));
                g1 = __longlong_as_double((long long)block_min_bits(lmin + 1 * 16, nwarps, lane));
                g2 = __longlong_as_double((long long)block_min_bits(lmin + 2 * 16, nwarps, lane));
                d12 = sqk[d];
                if (lane == 0) {
                    if (!(d12 < g1) && m1.v == g1) ra_new[r1] = (unsigned short)m1.a;
                    if (!(d12 < g2) && m2.v == g2) ra_new[r2] = (unsigned short)m2.a;
                }
            }

            AMX_T(5)
            // ---- Metropolis (anneal.rs:110-128) and the global best (:131-136), warp 0 only
            if (warp == 0) {
                double raw_new = raw_cur;
                if (moved) {
                    if (lane == 0) {
                        rm_new[r1] = fmin(g1, d12); if (d12 < g1) ra_new[r1] = (unsigned short)r2;
                        rm_new[r2] = fmin(g2, d12); if (d12 < g2) ra_new[r2] = (unsigned short)r1;
                    }
                    raw_new = fmin(fmin(fin_all, fmin(g1, d12)), fmin(g2, d12));
                }
                if (lane == 0) dirty_n[0] = 0;  // for the next iteration's list
                // Most moves leave the smallest distance where it was: raw_new == raw_cur bit for bit.  Then
                // new_metric == cur_metric, diff = +-0 and p = exp(-+0 / temp) is 1 for every temp > 0 (u < 1: accept)
                // and NaN for temp == 0 (0/0: reject) -- the same decision without the sqrt, the divide and the exp,
                // which are ~1,300 cycles of dependent latency on this one warp while fifteen others wait.
                double new_metric;
                bool accept;
                if (raw_new == raw_cur) {
                    new_metric = cur_metric;
                    accept = temp != 0.0;
                } else {
                    const double u = (double)w.w * 0x1p-32;  // anneal.rs:125
                    new_metric = sqrt(raw_new);
                    double diff = new_metric - cur_metric;
                    if (!p.minimize) diff *= -1.0;
                    const double p_acc = diff < 0.0 ? 1.0 : exp(-diff / temp);
                    accept = u < p_acc;
                }
                if (accept && moved && lane == 0) {
                    double v1 = cur[c * ns + r1], v2 = cur[c * ns + r2];
                    cur[c * ns + r1] = v2; cur[c * ns + r2] = v1;  // anneal.rs:34-39
                }
                if (accept) { cur_metric = new_metric; raw_cur = raw_new; }
                const bool improves = (p.minimize && cur_metric < global_best) || (!p.minimize && cur_metric > global_best);
                int best_mode = 0;
                if (accept) {
                    if (log_n < kLogCap) {
                        if (lane == 0) mlog[log_n] = make_ushort4((unsigned short)r1, (unsigned short)r2, (unsigned short)c, 0);
                        ++log_n;
                    } else {
                        overflow = true;
                    }
                }
                if (improves) {
                    best_mode = (materialized && !overflow) ? 1 : 2;
                    if (lane == 0) { dec->log_n = log_n; }
                    global_best = cur_metric;
                    materialized = true; overflow = false; log_n = 0;
                }
                if (lane == 0) { dec->accept = (accept && moved) ? 1 : 0; dec->best_mode = best_mode; }
                temp *= p.cooling;  // anneal.rs:139
            }
            AMX_T(6)
            __syncthreads();
            AMX_T(7)
            if (dec->accept) sel ^= 1;  // the tentative row minima become the committed ones
            const int best_mode = dec->best_mode;
            if (best_mode == 1) {
                const int ln = dec->log_n;
                for (int e = threadIdx.x; e < ln; e += blockDim.x) {
                    const ushort4 m = mlog[e];
                    best_out[(int)m.x * d + m.z] = cur[(int)m.z * ns + m.x];
                    best_out[(int)m.y * d + m.z] = cur[(int)m.z * ns + m.y];
                }
            } else if (best_mode == 2) {
                for (int e = threadIdx.x; e < nd; e += blockDim.x) {
                    int i = e / d, k = e - i * d;
                    best_out[e] = cur[k * ns + i];
                }
            }
        }
        if (threadIdx.x == 0) cta_best_commit(&cta_best, p, chain, global_best, materialized);
#if AMX_PROFILE
        if ((threadIdx.x == 0 || threadIdx.x == 200) && chain == 0)
            printf("tid %d, %llu steps, cycles per step: draws+decode %lld, stage+dirty list+barrier %lld, row pass %lld, warp minima+atomics %lld, barrier %lld, dirty rows+r1/r2 %lld, metropolis %lld, barrier %lld\n",
                   (int)threadIdx.x, p.iters, tsum[0] / (long long)p.iters, tsum[1] / (long long)p.iters, tsum[2] / (long long)p.iters, tsum[3] / (long long)p.iters,
                   tsum[4] / (long long)p.iters, tsum[5] / (long long)p.iters, tsum[6] / (long long)p.iters, tsum[7] / (long long)p.iters);
#endif
    }
}

// ---------------------------------------------------------------------------------------------------
// The same step with the row pass SCREENED on an 8-bit copy of the design (the default when it fits).
//
// The exact pass above spends 3 FP64 operations per dimension, row and target (the two moved rows and the
// ~2.6 dirty rows: 4.6 k cycles of a full FP64 pipe at n=1000, d=20) to find a handful of minima.  Almost all of
// those distances are nowhere near the minimum they compete for, and a coarse distance shows that.
//   * The chain keeps q[i][k] = rint((x[i][k] - colmin[k]) / s), s = (widest column range) / 255, four
//     dimensions to a 32-bit word, and nrm[i] = sum_k q[i][k]^2.  A swap exchanges two entries of one column,
//     so colmin and s hold for the whole chain and an accepted move exchanges two bytes.
//   * With D = sum_k (q[a][k] - q[b][k])^2 = nrm[a] + nrm[b] - 2 q[a].q[b] -- one IDP.4A per four dimensions,
//     exact integer arithmetic -- the triangle inequality gives | |x_a - x_b| / s - sqrt(D) | <= sqrt(d) (1 + 2e-12)
//     (every coordinate is within s (0.5 + 1e-12) of its quantised value).
//   * Every minimum the step needs has an upper bound U that costs nothing: a moved row r1' is still near r1's
//     old neighbour a, |x'_r1 - x_a|^2 = rowmin[r1] + (x'_r1[c] - x_a[c])^2 - (x_r1[c] - x_a[c])^2, likewise a dirty
//     row and the moved row it was nearest to, and a clean row j keeps rowmin[j].  Only a row whose D is at most
//     (sqrt(U)/s + sqrt(d))^2 can attain (or tie) such a minimum: those ~10 (row, target) items per step are
//     listed, and one thread each evaluates them in fp64 exactly as the pass above does -- full sums in the
//     reference's order.  Everything the chain decides on is still an exactly evaluated distance, so the
//     trajectory is the reference's bit for bit; the bound only decides what need not be evaluated, with
//     1e-5 of float slack on a margin that is 2 % wide.
//   * More than kMaxDirty dirty rows, a list longer than the block, or a design whose numbers cannot be quantised
//     (non-finite entries, a range that over- or underflows): the step is done by the exact pass.
// Per step: stage + bounds | integer pass | exact items | Metropolis, four barriers.
constexpr unsigned int kSlotCheck = 15u;   // item kind: does a moved row come closer to clean row j than rowmin[j]?
constexpr int kItemCap = 128;              // item slots of a step: 16 warps x kWarpItems (one thread each)
constexpr int kScreenDirty = 12;           // dirty rows the screened step takes (targets 2 .. 13; the exact pass takes any number)
constexpr int kScreenTargets = 2 + kScreenDirty;

// what the integer pass reads per target: thr - nrm[target] (D <= thr  <=>  nrm[j] - 2 q_j.q_t <= thrm, signed) and its row
struct ScreenTarget { int thrm; int row; };

// largest D that can belong to a squared distance <= U (U in the design's units), saturating at 2^31 - 1
__device__ __forceinline__ unsigned int screen_threshold(double U, float inv_s, float K) {
    float f = __double2float_ru(U > 0.0 ? U : 0.0);
    f = fmaxf(f, 1.1754944e-38f);  // no subnormals into the approximation (rounds up: still a bound)
    float sq;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(f));  // MUFU.SQRT, 2 ulp: covered by the 1e-5 below
    const float r = sq * inv_s * 1.00001f + K;
    const float t = r * r * 1.00001f + 1.0f;
    return t < 2.0e9f ? (unsigned int)t : 0x7FFFFFFFu;  // inf and NaN saturate too; every D is below 2^31
}

// upper bound of |x_i - x'_m|^2 where row m was i's nearest neighbour (squared distance rm) and moved in column c
// from x_old to x_new; xi = x[i][c]
__device__ __forceinline__ double moved_bound(double rm, double xi, double x_old, double x_new) {
    const double o = xi - x_old, nw = xi - x_new;
    const double o2 = o * o, n2 = nw * nw;
    return (rm + n2 - o2) + 1e-9 * (rm + n2 + o2);
}

// |a - x_j|^2 like sqdist_row (each square added in the reference's order), software-pipelined by hand: the loads of
// the next four dimensions are issued before the four dependent adds of the current ones, their squares after --
// a lone thread's chain is otherwise load latency + subtract + multiply + add per dimension, one after the other.
__device__ __forceinline__ double sqdist_chain(const double* a, int as, const double* xs, int ns, int j, int d) {
    const double* x = xs + j;
    double df = __dsub_rn(a[0], x[0]);
    double s = __dmul_rn(df, df);
    int k = 1;
    if (k + 4 <= d) {
        double q0, q1, q2, q3;
        {
            const double t0 = __dsub_rn(a[k * as], x[k * ns]), t1 = __dsub_rn(a[(k + 1) * as], x[(k + 1) * ns]);
            const double t2 = __dsub_rn(a[(k + 2) * as], x[(k + 2) * ns]), t3 = __dsub_rn(a[(k + 3) * as], x[(k + 3) * ns]);
            q0 = __dmul_rn(t0, t0); q1 = __dmul_rn(t1, t1); q2 = __dmul_rn(t2, t2); q3 = __dmul_rn(t3, t3);
        }
        k += 4;
#pragma unroll 1
        for (; k + 4 <= d; k += 4) {
            const double a0 = a[k * as], a1 = a[(k + 1) * as], a2 = a[(k + 2) * as], a3 = a[(k + 3) * as];
            const double x0 = x[k * ns], x1 = x[(k + 1) * ns], x2 = x[(k + 2) * ns], x3 = x[(k + 3) * ns];
            s = __dadd_rn(s, q0); s = __dadd_rn(s, q1); s = __dadd_rn(s, q2); s = __dadd_rn(s, q3);
            const double t0 = __dsub_rn(a0, x0), t1 = __dsub_rn(a1, x1), t2 = __dsub_rn(a2, x2), t3 = __dsub_rn(a3, x3);
            q0 = __dmul_rn(t0, t0); q1 = __dmul_rn(t1, t1); q2 = __dmul_rn(t2, t2); q3 = __dmul_rn(t3, t3);
        }
        s = __dadd_rn(s, q0); s = __dadd_rn(s, q1); s = __dadd_rn(s, q2); s = __dadd_rn(s, q3);
    }
#pragma unroll 1
    for (; k < d; ++k) { const double t = __dsub_rn(a[k * as], x[k * ns]); s = __dadd_rn(s, __dmul_rn(t, t)); }
    return s;
}

template <int DW>
__device__ __forceinline__ void load_words(const unsigned int* tw, unsigned int (&t)[DW > 0 ? DW : 1]) {  // tw 16-byte aligned
#pragma unroll
    for (int w = 0; w + 4 <= DW; w += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(tw + w);
        t[w] = v.x; t[w + 1] = v.y; t[w + 2] = v.z; t[w + 3] = v.w;
    }
    if ((DW & 3) >= 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(tw + (DW & ~3));
        t[DW & ~3] = v.x; t[(DW & ~3) + 1] = v.y;
    }
    if (DW & 1) t[DW - 1] = tw[DW - 1];
}

constexpr int kWarpItems = 8;  // items a warp can list per step (more: the step is done by the exact pass)

// Append `code` to the warp's own list for the lanes with `pred`; every lane of the warp calls it (no atomics: the
// count lives in a register of every lane).
__device__ __forceinline__ void warp_push(bool pred, unsigned int code, unsigned int* wl, int& cnt, int lane) {
    const unsigned int m = __ballot_sync(0xffffffffu, pred);
    if (m != 0u) {
        if (pred) {
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (pos < kWarpItems) wl[pos] = code;
        }
        cnt += __popc(m);
    }
}

// Integer pass of the screened step over the rows j outside {r1, r2}: which (row, target) distances can attain a
// minimum?  tqw[t][dwp]: the quantised targets (0, 1 the moved rows after the move, 2 + q dirty row q), tgt[t]
// their thresholds, tn0/tn1 the norms of the moved rows.  A thread walks two rows at a time so that a target's words
// are loaded once for both.  Items are (row | slot << 16).  DW > 0: words per row known at compile time; DW == 0: any.
// Returns the number of items the warp listed (more than kWarpItems: some were dropped).
template <int DW>
__device__ __forceinline__ int screen_row_pass(const unsigned int* qb, const unsigned int* nrm, const unsigned int* tqw, int n, int dw_rt, int dwp, int ns,
                                               int r1, int r2, int n_targets, const ScreenTarget* tgt, unsigned int tn0, unsigned int tn1,
                                               const double* rm, const unsigned short* ra, const unsigned int* tq, double* rm_new,
                                               unsigned short* ra_new, unsigned int* tq_new, unsigned int* wl, unsigned long long& allmin) {
    const int B = (int)blockDim.x, lane = threadIdx.x & 31;
    int cnt = 0;
    if (DW > 0) {
        for (int b0 = (int)threadIdx.x - lane; b0 < n; b0 += 2 * B) {  // warp-uniform trip count (warp_push votes)
            const bool in0 = b0 + lane < n, in1 = b0 + lane + B < n;
            const int j0 = in0 ? b0 + lane : 0, j1 = in1 ? b0 + lane + B : j0;
            const bool ok0 = in0 && j0 != r1 && j0 != r2, ok1 = in1 && j1 != r1 && j1 != r2;
            unsigned int x0[DW > 0 ? DW : 1], x1[DW > 0 ? DW : 1];
#pragma unroll
            for (int w = 0; w < DW; ++w) { x0[w] = qb[w * ns + j0]; x1[w] = qb[w * ns + j1]; }
            const int nj0 = (int)nrm[j0], nj1 = (int)nrm[j1];
            const int a0 = ra[j0], a1 = ra[j1];
            unsigned int dmin0 = 0xFFFFFFFFu, dmin1 = 0xFFFFFFFFu;  // min(D to r1', D to r2')
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                unsigned int tw[DW > 0 ? DW : 1];
                load_words<DW>(tqw + t * dwp, tw);
                const int thrm = tgt[t].thrm;
                unsigned int c0 = 0u, c1 = 0u;
#pragma unroll
                for (int w = 0; w < DW; ++w) { c0 = __dp4a(x0[w], tw[w], c0); c1 = __dp4a(x1[w], tw[w], c1); }
                const int v0 = nj0 - 2 * (int)c0, v1 = nj1 - 2 * (int)c1;
                const bool p0 = ok0 && v0 <= thrm, p1 = ok1 && v1 <= thrm;
                if (__any_sync(0xffffffffu, p0 || p1)) {
                    warp_push(p0, (unsigned int)j0 | ((unsigned int)t << 16), wl, cnt, lane);
                    warp_push(p1, (unsigned int)j1 | ((unsigned int)t << 16), wl, cnt, lane);
                }
                const unsigned int tn = t == 0 ? tn0 : tn1;
                dmin0 = min(dmin0, (unsigned int)v0 + tn); dmin1 = min(dmin1, (unsigned int)v1 + tn);
            }
            bool chk0 = false, chk1 = false;
            if (ok0 && a0 != r1 && a0 != r2) {  // clean row: the old minimum is still attained; can a moved row come closer?
                const double v = rm[j0];
                const unsigned int th = tq[j0];
                rm_new[j0] = v; ra_new[j0] = (unsigned short)a0; tq_new[j0] = th;
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
                allmin = vb < allmin ? vb : allmin;
                chk0 = dmin0 <= th;
            }
            if (ok1 && a1 != r1 && a1 != r2) {
                const double v = rm[j1];
                const unsigned int th = tq[j1];
                rm_new[j1] = v; ra_new[j1] = (unsigned short)a1; tq_new[j1] = th;
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
                allmin = vb < allmin ? vb : allmin;
                chk1 = dmin1 <= th;
            }
            if (__any_sync(0xffffffffu, chk0 || chk1)) {
                warp_push(chk0, (unsigned int)j0 | (kSlotCheck << 16), wl, cnt, lane);
                warp_push(chk1, (unsigned int)j1 | (kSlotCheck << 16), wl, cnt, lane);
            }
#pragma unroll 1
            for (int t = 2; t < n_targets; ++t) {  // the dirty row meets itself here (D = 0): its item stands for its distances to the moved rows
                unsigned int tw[DW > 0 ? DW : 1];
                load_words<DW>(tqw + t * dwp, tw);
                const int thrm = tgt[t].thrm;
                unsigned int c0 = 0u, c1 = 0u;
#pragma unroll
                for (int w = 0; w < DW; ++w) { c0 = __dp4a(x0[w], tw[w], c0); c1 = __dp4a(x1[w], tw[w], c1); }
                const bool p0 = ok0 && nj0 - 2 * (int)c0 <= thrm, p1 = ok1 && nj1 - 2 * (int)c1 <= thrm;
                if (__any_sync(0xffffffffu, p0 || p1)) {
                    warp_push(p0, (unsigned int)j0 | ((unsigned int)t << 16), wl, cnt, lane);
                    warp_push(p1, (unsigned int)j1 | ((unsigned int)t << 16), wl, cnt, lane);
                }
            }
        }
    } else {
        const int dw = dw_rt;
        for (int b0 = (int)threadIdx.x - lane; b0 < n; b0 += B) {
            const bool in = b0 + lane < n;
            const int j = in ? b0 + lane : 0;
            const bool ok = in && j != r1 && j != r2;
            const int nj = (int)nrm[j];
            unsigned int dmin = 0xFFFFFFFFu;
#pragma unroll 1
            for (int t = 0; t < n_targets; ++t) {
                unsigned int c = 0u;
                for (int w = 0; w < dw; ++w) c = __dp4a(qb[w * ns + j], tqw[t * dwp + w], c);
                const int v = nj - 2 * (int)c;
                warp_push(ok && v <= tgt[t].thrm, (unsigned int)j | ((unsigned int)t << 16), wl, cnt, lane);
                if (t < 2) dmin = min(dmin, (unsigned int)v + (t == 0 ? tn0 : tn1));
            }
            const int a = ra[j];
            bool chk = false;
            if (ok && a != r1 && a != r2) {
                const double v = rm[j];
                const unsigned int th = tq[j];
                rm_new[j] = v; ra_new[j] = (unsigned short)a; tq_new[j] = th;
                const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
                allmin = vb < allmin ? vb : allmin;
                chk = dmin <= th;
            }
            warp_push(chk, (unsigned int)j | (kSlotCheck << 16), wl, cnt, lane);
        }
    }
    return cnt;
}

// The exact pass of anneal_maximin_swap_kernel as a function: the screened kernel's way out for the steps its
// lists cannot hold.  lmin[1], lmin[2], lmin[3..] must hold +inf on entry; returns the minimum over the dirty rows.
__device__ __noinline__ double maximin_exact_step(const double* cur, const double* rowsx, int n, int d, int ns, int r1, int r2,
                                                  const double* rm, const unsigned short* ra, double* rm_new, unsigned short* ra_new,
                                                  const int* dirty, int nd_rows, unsigned long long* lmin, const double* sqk) {
    constexpr unsigned long long kInfBits = 0x7FF0000000000000ull;
    const int lane = threadIdx.x & 31;
    double fin = CUDART_INF;
    MinArg all{CUDART_INF, 0}, m1{CUDART_INF, 0}, m2{CUDART_INF, 0};
    for (int base = 0; base == 0 || base < nd_rows; base += kMaxDirty) {
        const int cnt = min(kMaxDirty, nd_rows - base);
        if (base != 0) {
            __syncthreads();
            if (threadIdx.x < kMaxDirty) lmin[3 + threadIdx.x] = kInfBits;
            __syncthreads();
        }
        MinArg dm[kMaxDirty];
#pragma unroll
        for (int q = 0; q < kMaxDirty; ++q) dm[q] = MinArg{CUDART_INF, 0};
        int di[kMaxDirty];
#pragma unroll
        for (int q = 0; q < kMaxDirty; ++q) di[q] = q < cnt ? dirty[base + q] : 0;
        const bool first = base == 0;
        // one instantiation for any count (unused slots walk row 0 and are ignored): this path is rare, its code size is not free
        maximin_row_pass<6>(cur, rowsx, n, d, ns, r1, r2, first, rm, ra, rm_new, ra_new, di, all, m1, m2, dm);
        if (first) {
            all = warp_minarg(all); m1 = warp_minarg(m1); m2 = warp_minarg(m2);
            if (lane == 0) {
                atomicMin(&lmin[0], (unsigned long long)__double_as_longlong(all.v));
                atomicMin(&lmin[1], (unsigned long long)__double_as_longlong(m1.v));
                atomicMin(&lmin[2], (unsigned long long)__double_as_longlong(m2.v));
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxDirty; ++q) {
            if (q < cnt) {
                dm[q] = warp_minarg(dm[q]);
                if (lane == 0) atomicMin(&lmin[3 + q], (unsigned long long)__double_as_longlong(dm[q].v));
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kMaxDirty; ++q) {
            if (q < cnt) {
                const unsigned long long g = lmin[3 + q];
                if (lane == 0 && (unsigned long long)__double_as_longlong(dm[q].v) == g) { rm_new[di[q]] = dm[q].v; ra_new[di[q]] = (unsigned short)dm[q].a; }
                fin = fmin(fin, __longlong_as_double((long long)g));
            }
        }
    }
    const double g1 = __longlong_as_double((long long)lmin[1]), g2 = __longlong_as_double((long long)lmin[2]), d12 = sqk[d];
    if (lane == 0) {
        if (!(d12 < g1) && m1.v == g1) ra_new[r1] = (unsigned short)m1.a;
        if (!(d12 < g2) && m2.v == g2) ra_new[r2] = (unsigned short)m2.a;
    }
    return fin;
}

template <int DW>
__global__ void __launch_bounds__(512) anneal_maximin_screen_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;
    const int dns = d * ns;
    const int dw = DW > 0 ? DW : (d + 3) >> 2;
    const int dwp = (dw + 3) & ~3;                            // words per staged target (16-byte rows)
    double* cur = smem_d;                                     // [k][i]
    double* rmin0 = cur + dns;                                // [2][ns]: committed / tentative row minima
    double* rowsx = rmin0 + 2 * ns;                           // [d][2]: the two moved rows AFTER the move
    double* sqk = rowsx + 2 * d;                              // [d] squared differences of the pair (r1, r2), then [d] = their sum
    // block minima (bit patterns) of the exact pass: [0] clean rows, [1] row r1, [2] row r2, [3 + q] dirty row q of a chunk
    unsigned long long* lmin = reinterpret_cast<unsigned long long*>(sqk + d + 1);  // [3 + kMaxDirty]
    unsigned long long* wmin = lmin + 3 + kMaxDirty;          // [16] per warp: smallest old minimum among its clean rows (bit pattern)
    unsigned long long* emin = wmin + 16;                     // [16] per evaluating warp: smallest listed distance of the step
    double* ival = reinterpret_cast<double*>(emin + 16);      // [kItemCap] the items' exact squared distances
    ScreenTarget* tgt = reinterpret_cast<ScreenTarget*>(ival + kItemCap);  // [kScreenTargets]
    Decision* dec = reinterpret_cast<Decision*>(tgt + kScreenTargets);
    // [kDrawBatch] decoded moves {r1, r2, c, accept draw}; aligned on the OFFSET (see the exact kernel)
    const size_t draws_off = ((size_t)(reinterpret_cast<unsigned char*>(dec + 1) - reinterpret_cast<unsigned char*>(smem_d)) + 15) & ~(size_t)15;
    uint4* draws = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(smem_d) + draws_off);
    unsigned int* tqw = reinterpret_cast<unsigned int*>(draws + kDrawBatch);  // [kScreenTargets][dwp] the quantised targets (16-byte rows)
    ushort4* mlog = reinterpret_cast<ushort4*>(tqw + kScreenTargets * dwp);
    unsigned int* qb = reinterpret_cast<unsigned int*>(mlog + kLogCap);       // [dw][ns] quantised design, 4 dimensions per word
    unsigned int* nrm = qb + dw * ns;                         // [ns] sum of the squared bytes of a row
    unsigned int* tq0 = nrm + ns;                             // [2][ns] largest D at which a moved row can come closer than the row's minimum
    unsigned int* tnrm = tq0 + 2 * ns;                        // [2] nrm of the two moved rows after the move
    int* dirty_n = reinterpret_cast<int*>(tnrm + 2);          // dirty_n[0] = count, then the list
    int* dirty = dirty_n + 1;                                 // [n] worst case
    int* wcnt = dirty + n;                                    // [16] items listed by each warp, [16] = some warp listed more than kWarpItems
    unsigned int* items = reinterpret_cast<unsigned int*>(wcnt + 17);   // [16][kWarpItems] (row | slot << 16), 0xFFFFFFFF: none
    unsigned short* iarg = reinterpret_cast<unsigned short*>(items + 16 * kWarpItems);  // [16][kWarpItems] who attains the item's distance
    unsigned short* rarg0 = iarg + 16 * kWarpItems;           // [2][ns] arguments of the row minima
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, last_warp = ((int)blockDim.x >> 5) - 1;

    const double* qp = p.qparams;  // colmin[d], s, 1/s, usable
    const double q_inv_s = qp[d + 1];
    const bool q_ok = qp[d + 2] != 0.0;
    const float inv_s_f = (float)q_inv_s * 1.00001f;
    const float K_f = sqrtf((float)d) * 1.00001f;
    const int nwarps = (int)blockDim.x >> 5, n_slots = nwarps * kWarpItems;

    __shared__ CtaBest cta_best;  // per-CTA fold of the chains' global bests (anneal.cuh)
    if (threadIdx.x == 0) cta_best_reset(&cta_best, p);
    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            const int i = e / d, k = e - i * d;
            cur[k * ns + i] = p.design[e];  // anneal.rs:89
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = p.rowmin0[i];
            rmin0[i] = v; rarg0[i] = p.rowarg0[i];
            tq0[i] = screen_threshold(v, inv_s_f, K_f);
        }
        if (q_ok) {
            for (int e = threadIdx.x; e < dw * n; e += blockDim.x) {
                const int i = e / dw, w = e - i * dw;
                unsigned int word = 0u;
                for (int b = 0; b < 4; ++b) {
                    const int k = 4 * w + b;
                    if (k < d) {
                        const double t = rint((p.design[(size_t)i * d + k] - qp[k]) * q_inv_s);
                        word |= (unsigned int)fmin(fmax(t, 0.0), 255.0) << (8 * b);
                    }
                }
                qb[w * ns + i] = word;
            }
        }
        if (threadIdx.x == 0) { dirty_n[0] = 0; wcnt[16] = 0; }
        if (threadIdx.x < 32) wmin[threadIdx.x] = 0x7FF0000000000000ull;  // wmin and emin: entries no warp writes stay +inf
        int sel = 0;  // which buffer holds the committed row minima
        double raw_cur = 0.0, cur_metric = 0.0, global_best = 0.0;
        double temp = p.t0;                                                             // anneal.rs:81
        bool materialized = false, overflow = false;
        int log_n = 0;
        if (warp == 0) {
            raw_cur = *p.raw0;
            cur_metric = sqrt(raw_cur);                                                 // :90
            global_best = cur_metric;                                                   // :94
        }
        double* best_out = cta_best_slot(p, &cta_best, nd);
        __syncthreads();
        if (q_ok) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                unsigned int s2 = 0u;
                for (int w = 0; w < dw; ++w) { const unsigned int x = qb[w * ns + i]; s2 = __dp4a(x, x, s2); }
                nrm[i] = s2;
            }
            __syncthreads();
        }

#if AMX_PROFILE
        long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
        unsigned long long n_items_sum = 0, n_exact_steps = 0;
#endif
        for (unsigned long long t = 0; t < p.iters; ++t) {
#if AMX_PROFILE
            const long long amx_t = chain == 0 ? (long long)t : -1;
#endif
            const unsigned int slot = (unsigned int)t & (kDrawBatch - 1);
            if (slot == 0) {
                for (int q = threadIdx.x; q < kDrawBatch; q += blockDim.x) {
                    const unsigned long long tq = t + (unsigned long long)q;
                    const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
                    // anneal.rs:29-31: r1, r2, c; the accept draw stays raw
                    draws[q] = make_uint4(mulhi_bound(w.x, (uint32_t)n), mulhi_bound(w.y, (uint32_t)n), mulhi_bound(w.z, (uint32_t)d), w.w);
                }
                __syncthreads();
            }
            const uint4 w = draws[slot];
            const int r1 = (int)w.x, r2 = (int)w.y, c = (int)w.z;
            const double* rm = rmin0 + sel * ns;
            const unsigned short* ra = rarg0 + sel * ns;
            const unsigned int* tq = tq0 + sel * ns;
            double* rm_new = rmin0 + (sel ^ 1) * ns;
            unsigned short* ra_new = rarg0 + (sel ^ 1) * ns;
            unsigned int* tq_new = tq0 + (sel ^ 1) * ns;
            const bool moved = r1 != r2;

            constexpr unsigned long long kInfBits = 0x7FF0000000000000ull;
            AMX_T(0)
            double raw_step = CUDART_INF, d12 = CUDART_INF;
            if (moved) {
                // ---- stage the two rows as they are AFTER the exchange (fp64 and quantised), list the dirty rows,
                //      bound every minimum of the step from above
                const double c1 = cur[c * ns + r1], c2 = cur[c * ns + r2];  // column c of r1, r2 before the exchange
                for (int k = threadIdx.x; k < d; k += blockDim.x) {
                    const double a = cur[k * ns + r1], b = cur[k * ns + r2];
                    const double a2 = (k == c) ? b : a, b2 = (k == c) ? a : b;
                    rowsx[2 * k] = a2;       // x'[r1][k]
                    rowsx[2 * k + 1] = b2;   // x'[r2][k]
                    const double df = __dsub_rn(a2, b2);
                    sqk[k] = __dmul_rn(df, df);
                }
                const int wc = c >> 2, sh = (c & 3) * 8;
                if (q_ok) {
                    for (int wv = (int)blockDim.x - 1 - (int)threadIdx.x; wv < dw; wv += blockDim.x) {  // from the last thread down: warp 0 decides, keep it light
                        unsigned int a = qb[wv * ns + r1], b = qb[wv * ns + r2];
                        if (wv == wc) { const unsigned int df = (a ^ b) & (0xFFu << sh); a ^= df; b ^= df; }
                        tqw[wv] = a; tqw[dwp + wv] = b;
                    }
                    for (int i = threadIdx.x; i < n; i += blockDim.x) {
                        const int a = ra[i];
                        if (i == r1 || i == r2) {
                            // the moved row is still near its old neighbour a (the pair (r1, r2) keeps its distance)
                            const bool first = i == r1;
                            const int other = first ? r2 : r1;
                            double U = rm[i];
                            if (a != other) U = moved_bound(U, cur[c * ns + a], first ? c1 : c2, first ? c2 : c1);
                            const unsigned int bo = (qb[wc * ns + i] >> sh) & 0xFFu, bn = (qb[wc * ns + other] >> sh) & 0xFFu;  // its byte in column c before / after
                            const unsigned int tn = nrm[i] + bn * bn - bo * bo;
                            tnrm[first ? 0 : 1] = tn;
                            tgt[first ? 0 : 1] = ScreenTarget{(int)screen_threshold(U, inv_s_f, K_f) - (int)tn, i};
                        } else if (a == r1 || a == r2) {
                            const int pos = atomicAdd(dirty_n, 1);  // dirty_n was zeroed before the last barrier
                            dirty[pos] = i;
                            if (pos < kScreenDirty) {  // target 2 + pos: the moved row it was nearest to is still near
                                const double U = moved_bound(rm[i], cur[c * ns + i], a == r1 ? c1 : c2, a == r1 ? c2 : c1);
                                tgt[2 + pos] = ScreenTarget{(int)screen_threshold(U, inv_s_f, K_f) - (int)nrm[i], i};
                                for (int wv = 0; wv < dw; ++wv) tqw[(2 + pos) * dwp + wv] = qb[wv * ns + i];
                            }
                        }
                    }
                } else {
                    for (int i = threadIdx.x; i < n; i += blockDim.x) {
                        const int a = ra[i];
                        if (i != r1 && i != r2 && (a == r1 || a == r2)) dirty[atomicAdd(dirty_n, 1)] = i;
                    }
                }
                __syncthreads();
                AMX_T(1)
                const int nd_rows = dirty_n[0];
                if (warp == last_warp) {
                    // the pair (r1, r2) itself: its squares summed in the reference's order (the last warp has the fewest rows)
                    double s12 = sqk[0];
#pragma unroll 4
                    for (int k = 1; k < d; ++k) s12 = __dadd_rn(s12, sqk[k]);
                    if (lane == 0) {
                        sqk[d] = s12;
                        // rows r1 and r2 unless an item comes at least as close (the exact pass writes them itself)
                        const unsigned int th = screen_threshold(s12, inv_s_f, K_f);
                        rm_new[r1] = s12; ra_new[r1] = (unsigned short)r2; tq_new[r1] = th;
                        rm_new[r2] = s12; ra_new[r2] = (unsigned short)r1; tq_new[r2] = th;
                    }
                }
                bool exact = !q_ok || nd_rows > kScreenDirty;
                if (!exact) {
                    // ---- integer pass over the rows: who can attain a minimum?
                    unsigned long long allmin = kInfBits;
                    const int cnt = screen_row_pass<DW>(qb, nrm, tqw, n, dw, dwp, ns, r1, r2, 2 + nd_rows, tgt, tnrm[0], tnrm[1], rm, ra, tq, rm_new, ra_new, tq_new,
                                                        items + warp * kWarpItems, allmin);
                    // ---- the distances this warp listed, in full and at once (no barrier: the list is the warp's own): two lanes per
                    //      item -- a clean row's item (and a dirty row meeting itself) stands for the distances to BOTH moved rows,
                    //      one each; every other item is one distance (the odd lane idles).  One code path for all items.
                    __syncwarp();
                    {
                        const int e = warp * kWarpItems + (lane >> 1), h = lane & 1;
                        const bool live = (lane >> 1) < min(cnt, kWarpItems);
                        const unsigned int it = live ? items[e] : 0xFFFFFFFFu;
                        const int j = (int)(it & 0xFFFFu);
                        const unsigned int sl = it >> 16;
                        const int trow = live && sl >= 2u && sl != kSlotCheck ? dirty[sl - 2u] : -1;
                        const bool both = live && (sl == kSlotCheck || trow == j);
                        double v = CUDART_INF;
                        if (live && (both || h == 0)) {
                            const bool mv = both || sl < 2u;  // against a moved row (staged) or against a dirty row (in place)
                            const double* a = mv ? rowsx + (both ? (unsigned int)h : sl) : cur + trow;
                            v = sqdist_chain(a, mv ? 2 : ns, cur, ns, j, d);
                        }
                        const double vo = __shfl_xor_sync(0xffffffffu, v, 1);
                        {   // the warp's share of the step's criterion: its clean rows' old minima and every distance it evaluated
                            const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
                            const unsigned long long m = vb < allmin ? vb : allmin;
                            const unsigned int hi = (unsigned int)(m >> 32), lo = (unsigned int)m;
                            const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
                            const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
                            if (lane == 0) {
                                wmin[warp] = ((unsigned long long)mhi << 32) | mlo;
                                if (cnt > kWarpItems) wcnt[16] = 1;
                            }
                        }
                        if (h == 0 && lane < 2 * kWarpItems) {
                            const double vm = both && vo < v ? vo : v;
                            const int arg = both ? (vo < v ? r2 : r1) : j;
                            if (live && sl == kSlotCheck) {  // the only item of its row: decided here
                                if (vm < rm[j]) { rm_new[j] = vm; ra_new[j] = (unsigned short)arg; tq_new[j] = screen_threshold(vm, inv_s_f, K_f); }
                            } else if (live) {
                                iarg[e] = (unsigned short)arg;
                            } else {
                                items[e] = 0xFFFFFFFFu;
                            }
                            ival[e] = vm;
                        }
#if AMX_PROFILE
                        if (lane == 0) wcnt[warp] = cnt;
#endif
                    }
                    AMX_T(2)
                    __syncthreads();
                    AMX_T(3)
                    exact = wcnt[16] != 0;
                    if (!exact) {
                        AMX_T(4)
                        AMX_T(5)
                        d12 = sqk[d];
                        if (warp == 0) {
                            // the step's criterion: every listed distance is a distance of the new design, the clean rows keep
                            // their old minima, the pair (r1, r2) its distance
                            // wmin[0..15] and emin[0..15] are one array: a load per lane (unused entries hold +inf)
                            unsigned long long m = wmin[lane];
                            const unsigned int hi = (unsigned int)(m >> 32), lo = (unsigned int)m;
                            const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
                            const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
                            raw_step = fmin(__longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo)), d12);
                        }
                        // who attains the minimum of a target: one warp per target, from the last warp down
#pragma unroll 1
                        for (int sl = last_warp - warp; sl < 2 + nd_rows; sl += nwarps) {
                            unsigned long long m = kInfBits;
                            int mf = 0x7FFFFFFF;
#pragma unroll 1
                            for (int f = lane; f < n_slots; f += 32) {
                                const unsigned long long vb = (unsigned long long)__double_as_longlong(ival[f]);
                                if ((items[f] >> 16) == (unsigned int)sl && vb < m) { m = vb; mf = f; }
                            }
                            const unsigned int hi = (unsigned int)(m >> 32), lo = (unsigned int)m;
                            const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
                            const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
                            const unsigned int f = __reduce_min_sync(0xffffffffu, (hi == mhi && lo == mlo) ? (unsigned int)mf : 0x7FFFFFFFu);
                            if (lane == 0 && f != 0x7FFFFFFFu) {
                                const double vmin = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
                                int row = -1;
                                if (sl < 2) { if (!(d12 < vmin)) row = sl == 0 ? r1 : r2; }
                                else row = dirty[sl - 2];
                                if (row >= 0) { rm_new[row] = vmin; ra_new[row] = iarg[f]; tq_new[row] = screen_threshold(vmin, inv_s_f, K_f); }
                            }
                        }
#if AMX_PROFILE
                        for (int q = 0; q < nwarps; ++q) n_items_sum += (unsigned long long)wcnt[q];
#endif
                    }
                }
                if (exact) {  // block-uniform
                    __syncthreads();
                    if (threadIdx.x < 3 + kMaxDirty) lmin[threadIdx.x] = kInfBits;
                    __syncthreads();
                    double fin = maximin_exact_step(cur, rowsx, n, d, ns, r1, r2, rm, ra, rm_new, ra_new, dirty, nd_rows, lmin, sqk);
                    fin = fmin(fin, __longlong_as_double((long long)lmin[0]));
                    const double g1 = __longlong_as_double((long long)lmin[1]), g2 = __longlong_as_double((long long)lmin[2]);
                    d12 = sqk[d];
                    raw_step = fmin(fmin(fin, fmin(g1, d12)), fmin(g2, d12));
                    if (threadIdx.x == 0) {
                        rm_new[r1] = fmin(g1, d12); if (d12 < g1) ra_new[r1] = (unsigned short)r2;
                        rm_new[r2] = fmin(g2, d12); if (d12 < g2) ra_new[r2] = (unsigned short)r1;
                    }
                    __syncthreads();
                    for (int i = threadIdx.x; i < n; i += blockDim.x) tq_new[i] = screen_threshold(rm_new[i], inv_s_f, K_f);
#if AMX_PROFILE
                    ++n_exact_steps;
#endif
                }
            }

            // ---- Metropolis (anneal.rs:110-128) and the global best (:131-136), warp 0 only
            if (warp == 0) {
                const double raw_new = moved ? raw_step : raw_cur;
                if (lane == 0) { dirty_n[0] = 0; wcnt[16] = 0; }  // for the next iteration's lists
                // raw_new == raw_cur bit for bit: same decision without the sqrt, the divide and the exp (see the exact kernel)
                double new_metric;
                bool accept;
                if (raw_new == raw_cur) {
                    new_metric = cur_metric;
                    accept = temp != 0.0;
                } else {
                    const double u = (double)w.w * 0x1p-32;
                    new_metric = sqrt(raw_new);
                    double diff = new_metric - cur_metric;
                    if (!p.minimize) diff *= -1.0;
                    const double p_acc = diff < 0.0 ? 1.0 : exp(-diff / temp);
                    accept = u < p_acc;
                }
                if (accept && moved && lane == 0) {
                    double v1 = cur[c * ns + r1], v2 = cur[c * ns + r2];
                    cur[c * ns + r1] = v2; cur[c * ns + r2] = v1;  // anneal.rs:34-39
                    if (q_ok) {
                        const int wc = c >> 2;
                        qb[wc * ns + r1] = tqw[wc]; qb[wc * ns + r2] = tqw[dwp + wc];
                        nrm[r1] = tnrm[0]; nrm[r2] = tnrm[1];
                    }
                }
                if (accept) { cur_metric = new_metric; raw_cur = raw_new; }
                const bool improves = (p.minimize && cur_metric < global_best) || (!p.minimize && cur_metric > global_best);
                int best_mode = 0;
                if (accept) {
                    if (log_n < kLogCap) {
                        if (lane == 0) mlog[log_n] = make_ushort4((unsigned short)r1, (unsigned short)r2, (unsigned short)c, 0);
                        ++log_n;
                    } else {
                        overflow = true;
                    }
                }
                if (improves) {
                    best_mode = (materialized && !overflow) ? 1 : 2;
                    if (lane == 0) { dec->log_n = log_n; }
                    global_best = cur_metric;
                    materialized = true; overflow = false; log_n = 0;
                }
                if (lane == 0) { dec->accept = (accept && moved) ? 1 : 0; dec->best_mode = best_mode; }
                temp *= p.cooling;  // anneal.rs:139
            }
            AMX_T(6)
            __syncthreads();
            AMX_T(7)
            if (dec->accept) sel ^= 1;  // the tentative row minima become the committed ones
            const int best_mode = dec->best_mode;
            if (best_mode == 1) {
                const int ln = dec->log_n;
                for (int e = threadIdx.x; e < ln; e += blockDim.x) {
                    const ushort4 m = mlog[e];
                    best_out[(int)m.x * d + m.z] = cur[(int)m.z * ns + m.x];
                    best_out[(int)m.y * d + m.z] = cur[(int)m.z * ns + m.y];
                }
            } else if (best_mode == 2) {
                for (int e = threadIdx.x; e < nd; e += blockDim.x) {
                    int i = e / d, k = e - i * d;
                    best_out[e] = cur[k * ns + i];
                }
            }
        }
        if (threadIdx.x == 0) cta_best_commit(&cta_best, p, chain, global_best, materialized);
#if AMX_PROFILE
        if ((threadIdx.x == 0 || threadIdx.x == 200 || threadIdx.x == blockDim.x - 1) && chain == 0)
            printf("screen tid %d, %llu steps, cycles per step: draws+decode %lld, stage+lists+barrier %lld, integer pass %lld, barrier %lld, exact items %lld, barrier %lld, args+metropolis %lld, barrier %lld; items per step %.2f, exact steps %llu\n",
                   (int)threadIdx.x, p.iters, tsum[0] / (long long)p.iters, tsum[1] / (long long)p.iters, tsum[2] / (long long)p.iters, tsum[3] / (long long)p.iters,
                   tsum[4] / (long long)p.iters, tsum[5] / (long long)p.iters, tsum[6] / (long long)p.iters, tsum[7] / (long long)p.iters,
                   (double)n_items_sum / (double)p.iters, n_exact_steps);
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x == 0 && chain == 0 && p.iters >= 504) {
            const long long t0 = g_amx_trace[0];
            for (int st = 0; st < 4; ++st)
                for (int pr = 0; pr < 8; ++pr) {
                    printf("trace step %d probe %d:", 500 + st, pr);
                    for (int q = 0; q < ((int)blockDim.x >> 5); ++q) printf(" %lld", g_amx_trace[(st * 8 + pr) * 16 + q] - t0);
                    printf("\n");
                }
        }
#endif
    }
}

// Column minima and the quantisation step of the start design, once per call: qp = colmin[d], s, 1/s, usable.
// One warp per column, lanes over the rows.
__global__ void __launch_bounds__(1024) anneal_quant_kernel(const double* design, int n, int d, double* qp) {
    __shared__ double wrange[32];
    __shared__ int wbad[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double rmax = 0.0;
    int bad = 0;
    for (int k = warp; k < d; k += nw) {
        double lo = CUDART_INF, hi = -CUDART_INF;
        for (int i = lane; i < n; i += 32) {
            const double v = design[(size_t)i * d + k];
            if (!(fabs(v) <= 1.7976931348623157e308)) bad = 1;
            lo = fmin(lo, v); hi = fmax(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) qp[k] = lo;
        rmax = fmax(rmax, hi - lo);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) { wrange[warp] = rmax; wbad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < nw; ++q) { rmax = fmax(rmax, wrange[q]); bad |= wbad[q]; }
        double s = rmax / 255.0;
        if (rmax == 0.0) s = 1.0;  // every column constant: all bytes 0, all errors 0
        const double inv_s = 1.0 / s;
        const bool ok = !bad && s >= 1e-30 && s <= 1e30 && inv_s >= 1e-30 && inv_s <= 1e30;
        qp[d] = s; qp[d + 1] = inv_s; qp[d + 2] = ok ? 1.0 : 0.0;
    }
}

// Nearest-neighbour table of the start design, once per call (every chain starts there): one warp
// per row i, lanes over j, distances in the reference's order.
__global__ void __launch_bounds__(256) anneal_rowmin0_kernel(const double* design, int n, int d, double* rowmin0, unsigned short* rowarg0) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const double* xi = design + (size_t)i * d;
    MinArg m{CUDART_INF, 0};
    for (int j = lane; j < n; j += 32) {
        if (j == i) continue;
        const double* xj = design + (size_t)j * d;
        double df = __dsub_rn(xi[0], xj[0]);
        double s = __dmul_rn(df, df);
        for (int k = 1; k < d; ++k) {
            df = __dsub_rn(xi[k], xj[k]);
            s = __dadd_rn(s, __dmul_rn(df, df));
        }
        m.take(s, j);
    }
    m = warp_minarg(m);
    if (lane == 0) { rowmin0[i] = m.v; rowarg0[i] = (unsigned short)m.a; }
}

size_t anneal_maximin_smem_bytes(unsigned long long n, unsigned long long d) {
    const unsigned long long ns = (n + 1) & ~1ull;
    size_t b = (size_t)(ns * d + 2 * ns + 2 * d + d + 1 + (3 + kMaxDirty) * 16) * sizeof(double);  // cur, rmin x2, rowsx, sqk, lmin
    b += sizeof(Decision) + (size_t)(1 + n) * sizeof(int) + 16;      // dec, dirty list, alignment slack
    b += (size_t)kDrawBatch * sizeof(uint4) + (size_t)kLogCap * sizeof(ushort4);
    b += (size_t)2 * ns * sizeof(unsigned short);
    return b + 64;
}

// the screened kernel: + quantised design, row norms, staged target words, thresholds, item list
static size_t anneal_maximin_screen_smem_bytes(unsigned long long n, unsigned long long d) {
    const unsigned long long ns = (n + 1) & ~1ull, dw = (d + 3) / 4, dwp = (dw + 3) & ~3ull;
    size_t b = (size_t)(ns * d + 2 * ns + 2 * d + d + 1 + 3 + kMaxDirty + 32 + kItemCap) * sizeof(double);  // cur, rmin x2, rowsx, sqk, lmin, wmin, emin, ival
    b += (size_t)kScreenTargets * sizeof(ScreenTarget) + sizeof(Decision) + (size_t)kDrawBatch * sizeof(uint4);
    b += (size_t)(kScreenTargets * dwp) * sizeof(unsigned int) + (size_t)kLogCap * sizeof(ushort4);
    b += (size_t)(dw * ns + ns + 2 * ns + 2 + 1 + n + 17 + 16 * kWarpItems) * sizeof(unsigned int);  // qb, nrm, tq x2, tnrm, dirty list, wcnt, items
    b += (size_t)(16 * kWarpItems + 2 * ns) * sizeof(unsigned short);                               // iarg, rarg x2
    return b + 64;
}

bool anneal_maximin_supported(unsigned long long n, unsigned long long d, int metric, int swap) {
    if (const char* e = cfg("MAXPRO_NO_ANNEAL_MAXIMIN_INC")) if (atoi(e) != 0) return false;
    return metric == 1 && swap != 0 && n >= 2 && n <= 65535 && d >= 1 && d <= 65535 && anneal_maximin_smem_bytes(n, d) <= kSmemBudget - kSmemSlack;
}

// The row pass screened on the 8-bit copy or evaluated in full.  The screened step replaces ~3 n d T FP64 operations by
// ~n (d/4) T integer ones plus a fixed ~3,000 cycles of lists, barriers and exact items, so it pays on wide designs only
// (one B200, 148 chains, us per step: (1000,20) 5.84 -> 4.27, (600,16) 4.37 -> 3.86, (500,20) 3.92 -> 3.54; (1000,10)
// 4.29 -> 4.40, (300,33) 4.17 -> 4.10, (700,5) 3.3 -> 4.6, (200,4) 2.3 -> 3.5 -- tools/anneal_maximin_ab.py).  MAXPRO_ANNEAL_MAXIMIN_SCREEN=1 forces it on every
// shape that has room for the copy (the parity tests do), =0 switches it off.
static bool anneal_maximin_screened(unsigned long long n, unsigned long long d) {
    if (d > 16384 || anneal_maximin_screen_smem_bytes(n, d) > kSmemBudget - kSmemSlack) return false;
    if (const char* e = cfg("MAXPRO_NO_ANNEAL_MAXIMIN_SCREEN")) if (atoi(e) != 0) return false;
    if (const char* e = cfg("MAXPRO_ANNEAL_MAXIMIN_SCREEN")) return atoi(e) != 0;
    return d >= 16 && n >= 500;
}

cudaError_t launch_anneal_maximin(const AnnealLaunch& L, int plan_sms, unsigned int* plan_grid) {
    const int n = (int)L.n, d = (int)L.d;
    const bool screened = anneal_maximin_screened(L.n, L.d);
    if (!plan_grid) {
        anneal_rowmin0_kernel<<<(n + 7) / 8, 256, 0, L.stream>>>(L.design, n, d, L.rowmin0, L.rowarg0);
        count_launch();
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
        if (screened) {
            if (!L.qparams) return cudaErrorInvalidValue;
            anneal_quant_kernel<<<1, 1024, 0, L.stream>>>(L.design, n, d, L.qparams);
            count_launch();
            if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
        }
    }
    AnnealParams p{};
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.0;
    p.minimize = L.minimize; p.swap = 1; p.resync = 0;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved; p.raw0 = L.raw0;
    p.cta_chain = L.cta_chain; p.cta_buf = L.cta_buf;
    p.rowmin0 = L.rowmin0; p.rowarg0 = L.rowarg0; p.qparams = L.qparams;
    const size_t smem = screened ? anneal_maximin_screen_smem_bytes(L.n, L.d) : anneal_maximin_smem_bytes(L.n, L.d);
    void (*kernel)(AnnealParams) = anneal_maximin_swap_kernel;
    if (screened) {
        switch ((L.d + 3) / 4) {  // words per quantised row: in registers up to eight
            case 1: kernel = anneal_maximin_screen_kernel<1>; break;
            case 2: kernel = anneal_maximin_screen_kernel<2>; break;
            case 3: kernel = anneal_maximin_screen_kernel<3>; break;
            case 4: kernel = anneal_maximin_screen_kernel<4>; break;
            case 5: kernel = anneal_maximin_screen_kernel<5>; break;
            case 6: kernel = anneal_maximin_screen_kernel<6>; break;
            case 7: kernel = anneal_maximin_screen_kernel<7>; break;
            case 8: kernel = anneal_maximin_screen_kernel<8>; break;
            default: kernel = anneal_maximin_screen_kernel<0>; break;
        }
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // Threads per chain.  One chain alone is fastest with one row per thread (up to 512 threads).  Many chains are not: a step
    // is barrier intervals of dependent shared-memory latency, which several small CTAs on one SM overlap and one large CTA
    // does not -- (500,10), 4,736 chains: 5.4e7 chain-steps/s with 512 threads, 8.1e7 with 256, 9.0e7 with 128; (300,6) 6.7e7
    // with 320, 1.98e8 with 64 (tools/anneal_many_chains.py).  So: the block size that puts the most CTAs on an SM, among
    // the CTAs the chains can fill, with at most six rows per thread; the larger block on a tie.
    int sms = plan_sms;
    if (sms <= 0) {
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    int threads = (int)((L.n + 31) / 32 * 32);
    if (threads > 512) threads = 512;
    if (const char* env = cfg("MAXPRO_ANNEAL_THREADS")) {
        const int v = atoi(env);
        if (v >= 32 && v <= 512 && v % 32 == 0) threads = v;
    } else if (L.n_chains > (unsigned long long)sms) {
        const unsigned long long needed = (L.n_chains + sms - 1) / sms;  // CTAs per SM the chains can fill
        const int t_max = threads;
        long long best_ctas = -1;
        for (int t = t_max; t >= 32; t -= 32) {
            if ((long long)t * 6 < (long long)L.n && t != t_max) break;  // more than six rows per thread
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, t, smem) != cudaSuccess) { cudaGetLastError(); continue; }
            const long long ctas = (long long)std::min<unsigned long long>((unsigned long long)per_sm, needed);
            if (ctas > best_ctas) { best_ctas = ctas; threads = t; }  // ">": the larger block on a tie (t descends)
        }
    }
    if (plan_grid) return anneal_resident_grid(kernel, threads, smem, plan_sms, L.n_chains, 1ull << 20, plan_grid);
    if (L.grid == 0 || L.grid > L.n_chains) return cudaErrorInvalidValue;
    kernel<<<L.grid, threads, smem, L.stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

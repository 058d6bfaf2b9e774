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

#include <cstdlib>

#ifndef AMX_PROFILE
#define AMX_PROFILE 0  // 1: clock64 sums per phase of the step, printed by thread 0 of CTA 0 (make EXTRA=-DAMX_PROFILE=1)
#endif
#if AMX_PROFILE
#include <cstdio>
#define AMX_T(i) { const long long now_ = clock64(); tsum[i] += now_ - tprev; tprev = now_; }
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

__global__ void __launch_bounds__(512) anneal_maximin_swap_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;
    const int dns = d * ns;
    double* cur = smem_d;                                     // [k][i]
    double* rmin0 = cur + dns;                                // [2][ns]: committed / tentative row minima
    double* rowsx = rmin0 + 2 * ns;                           // [d][2]: the two moved rows AFTER the move
    double* sqk = rowsx + 2 * d;                              // [d] squared differences of the pair (r1, r2), then [d] = their sum
    unsigned long long* lmin = reinterpret_cast<unsigned long long*>(sqk + d + 1);  // [3 + kMaxDirty] block minima (bit patterns)
    Decision* dec = reinterpret_cast<Decision*>(lmin + 3 + kMaxDirty);
    int* dirty_n = reinterpret_cast<int*>(dec + 1);           // dirty_n[0] = count, then the list
    int* dirty = dirty_n + 1;                                 // [n] worst case
    uint4* draws = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(dirty + n) + 15) & ~(uintptr_t)15);
    ushort4* mlog = reinterpret_cast<ushort4*>(draws + kDrawBatch);
    unsigned short* rarg0 = reinterpret_cast<unsigned short*>(mlog + kLogCap);  // [2][ns] their arguments
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

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
            const unsigned int slot = (unsigned int)t & (kDrawBatch - 1);
            if (slot == 0) {
                for (int q = threadIdx.x; q < kDrawBatch; q += blockDim.x) {
                    const unsigned long long tq = t + (unsigned long long)q;
                    const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
                    draws[q] = make_uint4(w.x, w.y, w.z, w.w);
                }
                __syncthreads();
            }
            const uint4 w = draws[slot];
            const int r1 = (int)mulhi_bound(w.x, (uint32_t)n);  // anneal.rs:29
            const int r2 = (int)mulhi_bound(w.y, (uint32_t)n);  // :30
            const int c = (int)mulhi_bound(w.z, (uint32_t)d);   // :31
            const double u = (double)w.w * 0x1p-32;
            const double* rm = rmin0 + sel * ns;
            const unsigned short* ra = rarg0 + sel * ns;
            double* rm_new = rmin0 + (sel ^ 1) * ns;
            unsigned short* ra_new = rarg0 + (sel ^ 1) * ns;
            const bool moved = r1 != r2;

            // Block minima go through 64-bit atomicMin on the bit patterns (squared distances are >= +0,
            // so the integer order is the numeric order): lmin[0] clean rows, [1] row r1, [2] row r2,
            // [3 + q] the dirty rows of the current chunk.  Who attains a minimum is written afterwards
            // by every warp whose partial equals it -- the decision needs the values only.
            constexpr unsigned long long kInfBits = 0x7FF0000000000000ull;
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
                if (threadIdx.x < 3 + kMaxDirty) lmin[threadIdx.x] = kInfBits;
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
                    if (base != 0) {  // more than kMaxDirty dirty rows: the previous chunk's minima have been read
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
                    AMX_T(3)
                    __syncthreads();
                    AMX_T(4)
                    // the dirty rows of this chunk: value and a neighbour that attains it
#pragma unroll
                    for (int q = 0; q < kMaxDirty; ++q) {
                        if (q < cnt) {
                            const unsigned long long g = lmin[3 + q];
                            if (lane == 0 && (unsigned long long)__double_as_longlong(dm[q].v) == g) { rm_new[di[q]] = dm[q].v; ra_new[di[q]] = (unsigned short)dm[q].a; }
                            fin_all = fmin(fin_all, __longlong_as_double((long long)g));
                        }
                    }
                }
                // rows r1 and r2: the minimum over the other rows and the pair (r1, r2) itself
                fin_all = fmin(fin_all, __longlong_as_double((long long)lmin[0]));
                g1 = __longlong_as_double((long long)lmin[1]);
                g2 = __longlong_as_double((long long)lmin[2]);
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
    size_t b = (size_t)(ns * d + 2 * ns + 2 * d + d + 1 + 3 + kMaxDirty) * sizeof(double);  // cur, rmin x2, rowsx, sqk, lmin
    b += sizeof(Decision) + (size_t)(1 + n) * sizeof(int) + 16;      // dec, dirty list, alignment slack
    b += (size_t)kDrawBatch * sizeof(uint4) + (size_t)kLogCap * sizeof(ushort4);
    b += (size_t)2 * ns * sizeof(unsigned short);
    return b + 64;
}

bool anneal_maximin_supported(unsigned long long n, unsigned long long d, int metric, int swap) {
    if (const char* e = cfg("MAXPRO_NO_ANNEAL_MAXIMIN_INC")) if (atoi(e) != 0) return false;
    return metric == 1 && swap != 0 && n >= 2 && n <= 65535 && d >= 1 && d <= 65535 && anneal_maximin_smem_bytes(n, d) <= kSmemBudget - kSmemSlack;
}

cudaError_t launch_anneal_maximin(const AnnealLaunch& L, int plan_sms, unsigned int* plan_grid) {
    const int n = (int)L.n, d = (int)L.d;
    if (!plan_grid) {
        anneal_rowmin0_kernel<<<(n + 7) / 8, 256, 0, L.stream>>>(L.design, n, d, L.rowmin0, L.rowarg0);
        count_launch();
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
    }
    AnnealParams p{};
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.0;
    p.minimize = L.minimize; p.swap = 1; p.resync = 0;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved; p.raw0 = L.raw0;
    p.cta_chain = L.cta_chain; p.cta_buf = L.cta_buf;
    p.rowmin0 = L.rowmin0; p.rowarg0 = L.rowarg0;
    const size_t smem = anneal_maximin_smem_bytes(L.n, L.d);
    int threads = (int)((L.n + 31) / 32 * 32);
    if (threads > 512) threads = 512;
    if (const char* e = cfg("MAXPRO_ANNEAL_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 512 && v % 32 == 0) threads = v; }
    cudaError_t e = cudaFuncSetAttribute(anneal_maximin_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (plan_grid) return anneal_resident_grid(anneal_maximin_swap_kernel, threads, smem, plan_sms, L.n_chains, 1ull << 20, plan_grid);
    if (L.grid == 0 || L.grid > L.n_chains) return cudaErrorInvalidValue;
    anneal_maximin_swap_kernel<<<L.grid, threads, smem, L.stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

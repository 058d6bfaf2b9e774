// K3, swap moves on the MaxPro criterion: the chain as a two-stage software pipeline.
//
// Same chain as anneal.cu (src/anneal.rs:23-40 swap_rows, :56-143 anneal_lhd), same draws,
// same decisions.  An iteration has a wide half -- the change of S over the 2(n-2) affected
// pairs, FP64-pipe bound -- and a narrow half that is one long dependent chain on a single
// warp: fold the partials, pow, subtract, divide, exp, compare (~1600 cycles).  Run back to
// back the SM idles through the narrow half.  Here the last warp of the CTA (the DECIDER)
// does the narrow half of iteration s while the other warps (the WORKERS) already compute the
// wide half of iteration s+1 -- before they know whether move s is accepted:
//
//   move s exchanges x[r1_s][c_s] and x[r2_s][c_s].  If rows(s+1) and rows(s) are disjoint,
//   the terms of iteration s+1 for partner rows j outside rows(s) read nothing the decision
//   can change, so they are summed from the design as it is (the decider may be writing the
//   two elements meanwhile; loads that see them are discarded).  The two partner rows in
//   rows(s) are evaluated twice -- as if move s were rejected and as if it were accepted --
//   from the staged copies of those rows (rows[] of iteration s, taken before the decision);
//   the decider adds the pair that matches its decision.
//
// The pipeline drains (one serial iteration) when consecutive moves share a row (~4/n of the
// iterations), at the periodic full re-sum, and at the first iteration.  One CTA barrier per
// pipelined iteration; the staging of a move's two rows is ordered by a named barrier among
// the workers and the warp that evaluates the four special terms.
//
// Both halves are latency chains on few warps, so their instruction counts are what is tuned:
//   - a move is decoded once, when its Philox block is drawn (128 iterations at a time):
//     {r1, r2, c, r1 == r2, -log u}.  With the threshold -log u at hand the acceptance test
//     u < exp(-diff/temp) (anneal.rs:118-128) becomes diff < temp * (-log u): no exp and no
//     divide on the decider's chain.  temp == 0 (underflow) accepts only diff < 0, as the
//     reference does (exp(-0/0) is NaN there and fails the comparison);
//   - when the change of S is small, |x| = |delta/S| < 2^-10, the criterion moves by
//     psi * expm1(log1p(x)/d), evaluated by two short series (relative error < 1e-16) instead
//     of a second pow: more accurate than the difference of two rounded powers and ~1/5 of
//     its latency.  psi is re-anchored by an exact pow at every full re-sum, and any larger
//     change takes the pow path.
#include "anneal.cuh"

#include <cstdlib>

namespace mp {

namespace {

constexpr int kRing = 2 * kDrawBatch;  // decoded moves of two batches: the running one and the next

// A decoded swap move (anneal.rs:29-31 + the acceptance draw of :125).
struct __align__(16) Move {
    int r1, r2, c;
    int noop;    // r1 == r2: the design does not change
    double thr;  // -log(u): accept when diff < temp * thr
    double pad;
};

__device__ __forceinline__ void fill_moves(Move* ring, unsigned long long first, int count, int threads, int n, int d,
                                           uint32_t key0, uint32_t key1) {
    for (int q = threadIdx.x; q < count; q += threads) {
        const unsigned long long tq = first + (unsigned long long)q;
        const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
        Move m;
        m.r1 = (int)mulhi_bound(w.x, (uint32_t)n);  // anneal.rs:29
        m.r2 = (int)mulhi_bound(w.y, (uint32_t)n);  // :30
        m.c = (int)mulhi_bound(w.z, (uint32_t)d);   // :31
        m.noop = m.r1 == m.r2;
        // u = w.w 2^-32 (:125).  u == 0 passes whenever exp(-diff/temp) > 0, i.e. diff/temp < 745.13...
        m.thr = w.w ? -log((double)w.w * 0x1p-32) : 745.13;
        m.pad = 0.0;
        ring[(unsigned int)tq & (kRing - 1)] = m;
    }
}

__device__ __forceinline__ Move load_move(const Move* ring, unsigned long long s) {
    return ring[(unsigned int)s & (kRing - 1)];
}

__device__ __forceinline__ void named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// The four special terms of an iteration on one warp, eight lanes each.  Term q: partner row
// r1 (q&1 == 0) or r2 of the previous move, as it is if that move is rejected (q>>1 == 0) or
// accepted (its c_prev element exchanged), against the rows of the next move.
__device__ __forceinline__ void special_terms(const double2* rows_prev, const double2* rows_next, int d, int c_prev,
                                              int c_next, double* out) {
    const int lane = threadIdx.x & 31, q = lane >> 3, sub = lane & 7;
    const int jidx = q & 1, applied = q >> 1;
    double A = 1.0, B = 1.0, uu = 0.0, vv = 0.0;
    for (int k = sub; k < d; k += 8) {
        const double2 rp = rows_prev[k];
        const bool other = applied && k == c_prev;
        const double xj = (jidx != 0) != other ? rp.y : rp.x;
        const double2 ab = rows_next[k];
        if (k == c_next) { uu = ab.x - xj; vv = ab.y - xj; }
        else { A *= ab.x - xj; B *= ab.y - xj; }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        A *= __shfl_xor_sync(0xffffffffu, A, o);
        B *= __shfl_xor_sync(0xffffffffu, B, o);
        uu += __shfl_xor_sync(0xffffffffu, uu, o);  // one lane holds the value, the others 0
        vv += __shfl_xor_sync(0xffffffffu, vv, o);
    }
    const double au = A * uu, av = A * vv, bu = B * uu, bv = B * vv;
    const double p1 = fma(au, au, kEps), p1n = fma(av, av, kEps);
    const double p2 = fma(bv, bv, kEps), p2n = fma(bu, bu, kEps);
    if (sub == 0) out[q] = (p1 - p1n) / (p1 * p1n) + (p2 - p2n) / (p2 * p2n);
}

// psi' - psi for S' = S (1 + x), |x| < 2^-10:  psi * expm1(log1p(x) / d)
__device__ __forceinline__ double small_metric_change(double psi, double x, double inv_d) {
    double l = fma(x, -1.0 / 6.0, 1.0 / 5.0);  // log1p(x) = x (1 - x/2 + x^2/3 - x^3/4 + x^4/5 - x^5/6)
    l = fma(x, l, -1.0 / 4.0);
    l = fma(x, l, 1.0 / 3.0);
    l = fma(x, l, -1.0 / 2.0);
    l = fma(x, l, 1.0);
    const double y = x * l * inv_d;
    double e = fma(y, 1.0 / 120.0, 1.0 / 24.0);  // expm1(y) = y (1 + y/2 + y^2/6 + y^3/24 + y^4/120)
    e = fma(y, e, 1.0 / 6.0);
    e = fma(y, e, 1.0 / 2.0);
    e = fma(y, e, 1.0);
    return psi * (y * e);
}

// Stage the two rows of a move for the wide half.  rp[kk], kk < 4G: {x[r1][k], x[r2][k]} with
// k = kk + (kk >= c), i.e. the untouched dimensions in order, padded with {1, 1} (the design
// carries three zero columns after the last dimension, so a padded slot multiplies by
// (1 - 0) = 1 exactly and the loop over groups of four needs no guards); rp[4G] is dimension c.
// rn[k]: the same rows in natural order, for the special terms of the NEXT iteration.
__device__ __forceinline__ void stage_rows(const double* cur, double2* rp, double2* rn, int d, int ns, int G,
                                           int r1, int r2, int c, int workers) {
    for (int kk = threadIdx.x; kk <= 4 * G; kk += workers) {
        const int k = kk == 4 * G ? c : kk + (kk >= c ? 1 : 0);
        rp[kk] = (kk < d - 1 || kk == 4 * G) ? make_double2(cur[k * ns + r1], cur[k * ns + r2]) : make_double2(1.0, 1.0);
    }
    for (int k = threadIdx.x; k < d; k += workers) rn[k] = make_double2(cur[k * ns + r1], cur[k * ns + r2]);
}

// The wide half: change of S over the partner rows of move (r1, r2, c), this thread's share.
// A thread owns PPT pairs of neighbouring rows (pair slots tid, tid + workers, ...: 16-byte
// loads, conflict-free across the lanes) and walks the d-1 untouched dimensions in groups of
// four, the loads of group g+1 in flight while group g is multiplied in.  The design is read
// exactly once per iteration (8 n d bytes of shared-memory traffic, the floor of this step).
template <int PPT>
__device__ __forceinline__ void bulk_load(double2 (&xv)[PPT][4], double2 (&rv)[4], const char* const (&bp)[PPT],
                                          const double2* rp, int g, int c, unsigned S) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int kk = 4 * g + t;
        const unsigned off = (unsigned)(kk + (kk >= c ? 1 : 0)) * S;
        rv[t] = rp[kk];
#pragma unroll
        for (int s = 0; s < PPT; ++s) xv[s][t] = *reinterpret_cast<const double2*>(bp[s] + off);
    }
}
template <int PPT>
__device__ __forceinline__ void bulk_mul(double (&A)[PPT][2], double (&B)[PPT][2], const double2 (&xv)[PPT][4], const double2 (&rv)[4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            A[s][0] *= rv[t].x - xv[s][t].x; B[s][0] *= rv[t].y - xv[s][t].x;
            A[s][1] *= rv[t].x - xv[s][t].y; B[s][1] *= rv[t].y - xv[s][t].y;
        }
}
template <int PPT>
__device__ __forceinline__ double pipe_bulk(const double* cur, const double2* rp, int n, int ns, int G, int c,
                                            int r1, int r2, int e1, int e2, int workers, bool in_cube) {
    const unsigned S = (unsigned)ns * 8u;
    double delta = 0.0;
    for (int p0 = threadIdx.x; 2 * p0 < n; p0 += PPT * workers) {
        int j0[PPT];
        const char* bp[PPT];
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            j0[s] = 2 * (p0 + s * workers);
            bp[s] = reinterpret_cast<const char*>(cur + min(j0[s], ns - 2));  // slots past the end are masked below
        }
        double A[PPT][2], B[PPT][2];
#pragma unroll
        for (int s = 0; s < PPT; ++s) { A[s][0] = A[s][1] = B[s][0] = B[s][1] = 1.0; }
        // two register sets in rotation: group g+1 is loaded (the last one twice) before group g is used
        double2 xa[PPT][4], xb[PPT][4], ra[4], rb[4];
        if (G > 0) bulk_load<PPT>(xa, ra, bp, rp, 0, c, S);
#pragma unroll 2
        for (int g = 0; g < G; ++g) {
            bulk_load<PPT>(xb, rb, bp, rp, min(g + 1, G - 1), c, S);
            bulk_mul<PPT>(A, B, xa, ra);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                ra[t] = rb[t];
#pragma unroll
                for (int s = 0; s < PPT; ++s) xa[s][t] = xb[s][t];
            }
        }
        const double2 ab = rp[4 * G];
        double num = 0.0, den = 1.0;
#pragma unroll
        for (int s = 0; s < PPT; ++s) {
            const double2 xc = *reinterpret_cast<const double2*>(bp[s] + (unsigned)c * S);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0[s] + h;
                const double xjc = h ? xc.y : xc.x;
                const bool live = j != r1 && j != r2 && j != e1 && j != e2 && j < n;
                const double vv = ab.y - xjc;
                double uu = ab.x - xjc;
                if (in_cube && !live) uu = vv;  // p1 == p1', p2 == p2': the row contributes exactly 0
                const double au = A[s][h] * uu, av = A[s][h] * vv, bu = B[s][h] * uu, bv = B[s][h] * vv;
                const double p1 = fma(au, au, kEps), p1n = fma(av, av, kEps);
                const double p2 = fma(bv, bv, kEps), p2n = fma(bu, bu, kEps);
                if (in_cube) {
                    // (p1-p1')/(p1 p1') + (p2-p2')/(p2 p2') over one denominator (>= 1e-48 in the unit cube),
                    // folded into the thread's running fraction
                    const double d1 = p1 * p1n, d2 = p2 * p2n;
                    const double nm = fma(p1 - p1n, d2, (p2 - p2n) * d1);
                    const double dn = d1 * d2;
                    num = fma(num, dn, nm * den);
                    den *= dn;
                } else if (live) {
                    delta += (p1 - p1n) / (p1 * p1n) + (p2 - p2n) / (p2 * p2n);
                }
            }
        }
        if (in_cube) delta += num * fast_rcp(den);  // den >= 1e-192 for PPT = 2
    }
    return delta;
}

template <int PPT>
__global__ void __launch_bounds__(PPT == 2 ? 320 : 480) anneal_swap_pipe_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;
    const int dns = d * ns;
    const int G = (d + 2) / 4;                              // groups of four untouched dimensions, (d-1) rounded up
    double* cur = smem_d;                                   // [k][i], then three zero columns
    double* part = smem_d + dns + 3 * ns;                   // [2][32] warp partials, by iteration parity
    double* spec = part + 64;                               // [2][4] special terms {rejected r1, r2, accepted r1, r2}
    Decision* decs = reinterpret_cast<Decision*>(spec + 8); // [2] by iteration parity
    Move* ring = reinterpret_cast<Move*>(spec + 12);        // kRing decoded moves
    double2* rows = reinterpret_cast<double2*>(ring + kRing);   // [2][d] {x[r1][k], x[r2][k]}, by iteration parity
    double2* rowsp = rows + 2 * d;                              // [2][d+3] the same in bulk order (stage_rows)
    ushort4* mlog = reinterpret_cast<ushort4*>(rowsp + 2 * (d + 3));  // kLogCap accepted moves {r1, r2, c}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5, wwarps = nwarps - 2, workers = wwarps * 32;
    const bool worker = warp < wwarps, decider = warp == wwarps, specialist = warp == wwarps + 1;

    __shared__ CtaBest cta_best;  // per-CTA fold of the chains' global bests (anneal.cuh)
    if (threadIdx.x == 0) cta_best_reset(&cta_best, p);
    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        if (ns != n)
            for (int k = threadIdx.x; k < d; k += blockDim.x) cur[k * ns + n] = 0.0;
        for (int e = threadIdx.x; e < 3 * ns; e += blockDim.x) cur[dns + e] = 0.0;
        int ok = 1;
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            const int i = e / d, k = e - i * d;
            const double v = p.design[e];
            ok &= (v >= 0.0 && v <= 1.0) ? 1 : 0;
            cur[k * ns + i] = v;  // anneal.rs:89
        }
        fill_moves(ring, 0ull, kRing, blockDim.x, n, d, key0, key1);
        const bool in_cube = __syncthreads_and(ok) != 0;
        // chain state, kept by the decider warp (every lane holds the same values)
        double raw_cur = 0.0, cur_metric = 0.0, global_best = 0.0;
        double temp = p.t0;                                                             // anneal.rs:81
        bool materialized = false, overflow = false;
        int log_n = 0, prev_applied = 0;
        if (decider) {
            raw_cur = *p.raw0;  // S of the start design, computed once for all chains (anneal_raw0_kernel)
            cur_metric = metric_from_raw<true>(raw_cur, n, p.n_pairs, p.inv_d);         // :90
            global_best = cur_metric;                                                   // :94
        }
        double* best_out = cta_best_slot(p, &cta_best, nd);
        unsigned long long until_resync = p.resync;
        bool have_bulk = false, use_fix = false;

        for (unsigned long long s = 0; s < p.iters; ++s) {
            const int cb = (int)(s & 1), nb = cb ^ 1;
            const Move m = load_move(ring, s);
            const bool full = --until_resync == 0;  // periodic re-sum: apply, score everything, maybe undo
            if (full) until_resync = p.resync;
            if (!have_bulk) {
                // ---- drained pipeline: the wide half of iteration s on the design as it stands
                if (full) {
                    if (threadIdx.x == 0) {
                        double v1 = cur[m.c * ns + m.r1], v2 = cur[m.c * ns + m.r2];
                        cur[m.c * ns + m.r1] = v2; cur[m.c * ns + m.r2] = v1;  // anneal.rs:34-39
                    }
                    __syncthreads();
                    publish_partial<true>(full_score_partial<true, 4>(cur, n, d, ns, in_cube), part + 32 * cb);
                } else if (!m.noop && worker) {
                    stage_rows(cur, rowsp + cb * (d + 3), rows + cb * d, d, ns, G, m.r1, m.r2, m.c, workers);
                    named_sync(1, workers);
                    publish_partial<true>(pipe_bulk<PPT>(cur, rowsp + cb * (d + 3), n, ns, G, m.c, m.r1, m.r2, -1, -1, workers, in_cube),
                                          part + 32 * cb);
                }
                __syncthreads();
            }
            // ---- may the workers run ahead?  Not past a re-sum, not when the next move shares a row
            Move mn = m;
            bool next_ok = false;
            if (s + 1 < p.iters) {
                mn = load_move(ring, s + 1);
                const bool shares_row = !m.noop && !mn.noop && (mn.r1 == m.r1 || mn.r1 == m.r2 || mn.r2 == m.r1 || mn.r2 == m.r2);
                next_ok = !full && until_resync != 1 && !shares_row;
            }
            const bool run_ahead = next_ok && !mn.noop;
            if (worker) {
                if (((unsigned int)s & (kDrawBatch - 1)) == 0 && s != 0)
                    fill_moves(ring, s + kDrawBatch, kDrawBatch, workers, n, d, key0, key1);  // the half nobody reads any more
                if (run_ahead) {
                    stage_rows(cur, rowsp + nb * (d + 3), rows + nb * d, d, ns, G, mn.r1, mn.r2, mn.c, workers);
                    named_sync(2, workers + 32);  // with the specialist warp
                    const int e1 = m.noop ? -1 : m.r1, e2 = m.noop ? -1 : m.r2;
                    publish_partial<true>(pipe_bulk<PPT>(cur, rowsp + nb * (d + 3), n, ns, G, mn.c, mn.r1, mn.r2, e1, e2, workers, in_cube),
                                          part + 32 * nb);
                }
            } else if (specialist) {
                if (run_ahead) {
                    named_sync(2, workers + 32);
                    // the two rows of move s under both outcomes (staged copies: the decider may be exchanging the originals)
                    if (!m.noop) special_terms(rows + cb * d, rows + nb * d, d, m.c, mn.c, spec + nb * 4);
                }
            } else {
                // ---- Metropolis (anneal.rs:110-128) and the global best (:131-136)
                double raw_new, new_metric, diff;
                bool exact = true;
                if (full) {
                    raw_new = warp_total<true>(part + 32 * cb, nwarps);
                } else {
                    double delta = 0.0;
                    if (!m.noop) {
                        delta = warp_total<true>(part + 32 * cb, wwarps);
                        if (use_fix) delta += spec[cb * 4 + 2 * prev_applied] + spec[cb * 4 + 2 * prev_applied + 1];
                    }
                    raw_new = raw_cur + delta;
                    const double x = delta / raw_cur;
                    if (fabs(x) < 0x1p-10 && n >= 2) {
                        exact = false;
                        diff = small_metric_change(cur_metric, x, p.inv_d);
                        new_metric = cur_metric + diff;
                    }
                }
                if (exact) {
                    new_metric = metric_from_raw<true>(raw_new, n, p.n_pairs, p.inv_d);
                    diff = new_metric - cur_metric;
                }
                if (!p.minimize) diff *= -1.0;
                // u < (diff < 0 ? 1 : exp(-diff/temp))  <=>  diff < 0 or diff < temp * (-log u)
                const bool accept = diff < 0.0 || diff < temp * m.thr;
                if (accept != full && lane == 0) {
                    // apply an accepted move / undo the rejected move of a re-sum iteration
                    double v1 = cur[m.c * ns + m.r1], v2 = cur[m.c * ns + m.r2];
                    cur[m.c * ns + m.r1] = v2; cur[m.c * ns + m.r2] = v1;
                }
                prev_applied = accept ? 1 : 0;
                if (accept) { cur_metric = new_metric; raw_cur = raw_new; }
                const bool improves = (p.minimize && cur_metric < global_best) || (!p.minimize && cur_metric > global_best);
                if (accept) {
                    if (log_n < kLogCap) {
                        if (lane == 0) mlog[log_n] = make_ushort4((unsigned short)m.r1, (unsigned short)m.r2, (unsigned short)m.c, 0);
                        ++log_n;
                    } else {
                        overflow = true;
                    }
                }
                int best_mode = 0;
                if (improves) {
                    __syncwarp();
                    if (materialized && !overflow) {
                        // the HBM best differs from cur only where a logged move touched it
                        for (int e = lane; e < log_n; e += 32) {
                            const ushort4 g = mlog[e];
                            best_out[(int)g.x * d + g.z] = cur[(int)g.z * ns + g.x];
                            best_out[(int)g.y * d + g.z] = cur[(int)g.z * ns + g.y];
                        }
                        __syncwarp();
                    } else {
                        best_mode = 2;  // everybody rewrites it after the barrier
                    }
                    global_best = cur_metric;
                    materialized = true; overflow = false; log_n = 0;
                }
                if (lane == 0) { decs[cb].accept = accept ? 1 : 0; decs[cb].best_mode = best_mode; }
                temp *= p.cooling;  // anneal.rs:139
            }
            __syncthreads();
            if (decs[cb].best_mode == 2) {
                for (int e = threadIdx.x; e < nd; e += blockDim.x) {
                    int i = e / d, k = e - i * d;
                    best_out[e] = cur[k * ns + i];
                }
                __syncthreads();  // the next decision may write cur
            }
            have_bulk = next_ok;
            use_fix = next_ok && !m.noop;
        }
        if (decider && lane == 0) cta_best_commit(&cta_best, p, chain, global_best, materialized);
    }
}

size_t pipe_smem_bytes(unsigned long long n, unsigned long long d) {
    const unsigned long long ns = (n + 1) & ~1ull;
    return (size_t)(ns * (d + 3) + 64 + 8 + 4 + 4 * d + 4 * (d + 3)) * sizeof(double) + (size_t)kRing * sizeof(Move) + (size_t)kLogCap * sizeof(ushort4);
}

}  // namespace

bool anneal_pipe_supported(unsigned long long n, unsigned long long d, int metric, int swap) {
    if (const char* e = cfg("MAXPRO_NO_ANNEAL_PIPE")) if (atoi(e) != 0) return false;
    return metric == 0 && swap && n >= 4 && n <= 65535 && d >= 1 && d <= 65535 && pipe_smem_bytes(n, d) <= kSmemBudget - kSmemSlack;
}

cudaError_t launch_anneal_pipe(const AnnealLaunch& L, int plan_sms, unsigned int* plan_grid) {
    AnnealParams p;
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.01 / (double)L.n;
    p.minimize = L.minimize; p.swap = 1;
    unsigned long long rs = 16 * L.n;
    p.resync = rs < 1024 ? 1024 : rs;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved; p.raw0 = L.raw0;
    p.cta_chain = L.cta_chain; p.cta_buf = L.cta_buf;
    const size_t smem = pipe_smem_bytes(L.n, L.d);
    // worker warps: one thread per PPT pairs of partner rows; plus the decider and the specialist warp
    const int ppt = L.n >= 768 ? 2 : 1;
    const unsigned long long slots = ((L.n + 1) / 2 + ppt - 1) / ppt;
    int wwarps = (int)((slots + 31) / 32);
    const int cap = ppt == 2 ? 8 : 13;
    if (wwarps > cap) wwarps = cap;
    if (const char* e = cfg("MAXPRO_ANNEAL_THREADS")) { int v = atoi(e) / 32; if (v >= 1 && v <= cap) wwarps = v; }
    const int threads = (wwarps + 2) * 32;
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        if (plan_grid) { e = anneal_resident_grid(kernel, threads, smem, plan_sms, L.n_chains, 1ull << 20, plan_grid); return; }
        if (L.grid == 0 || L.grid > L.n_chains) { e = cudaErrorInvalidValue; return; }
        kernel<<<L.grid, threads, smem, L.stream>>>(p);
    };
    if (ppt == 2) go(anneal_swap_pipe_kernel<2>); else go(anneal_swap_pipe_kernel<1>);
    if (e != cudaSuccess || plan_grid) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

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
// the workers only.
#include "anneal.cuh"

#include <cstdlib>

namespace mp {

namespace {

constexpr int kRing = 2 * kDrawBatch;  // swap draws of two batches: the running one and the next

struct Move {
    int r1, r2, c;
    double u;
    bool noop;  // r1 == r2: the design does not change
};

__device__ __forceinline__ Move load_move(const uint4* draws, unsigned long long s, int n, int d) {
    const uint4 w = draws[(unsigned int)s & (kRing - 1)];
    Move m;
    m.r1 = (int)mulhi_bound(w.x, (uint32_t)n);  // anneal.rs:29
    m.r2 = (int)mulhi_bound(w.y, (uint32_t)n);  // :30
    m.c = (int)mulhi_bound(w.z, (uint32_t)d);   // :31
    m.u = (double)w.w * 0x1p-32;                // :125
    m.noop = m.r1 == m.r2;
    return m;
}

__device__ __forceinline__ void fill_draws(uint4* draws, unsigned long long first, int count, int workers,
                                           uint32_t key0, uint32_t key1) {
    for (int q = threadIdx.x; q < count; q += workers) {
        const unsigned long long tq = first + (unsigned long long)q;
        const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
        draws[(unsigned int)tq & (kRing - 1)] = make_uint4(w.x, w.y, w.z, w.w);
    }
}

__device__ __forceinline__ void workers_sync(int workers) { asm volatile("bar.sync 1, %0;" ::"r"(workers) : "memory"); }

// One partner row of move `next` that belongs to the previous move: jidx picks r1/r2 of the
// previous move, `applied` the version of that row (its c_prev element exchanged or not).
__device__ __forceinline__ double special_term(const double2* rows_prev, const double2* rows_next, int d, int c_prev,
                                               int c_next, int jidx, int applied) {
    double A = 1.0, B = 1.0, uu = 0.0, vv = 0.0;
    for (int k = 0; k < d; ++k) {
        const double2 rp = rows_prev[k];
        const bool other = applied && k == c_prev;
        const double xj = (jidx != 0) != other ? rp.y : rp.x;
        const double2 ab = rows_next[k];
        if (k == c_next) { uu = ab.x - xj; vv = ab.y - xj; }
        else { A *= ab.x - xj; B *= ab.y - xj; }
    }
    const double au = A * uu, av = A * vv, bu = B * uu, bv = B * vv;
    const double p1 = fma(au, au, kEps), p1n = fma(av, av, kEps);
    const double p2 = fma(bv, bv, kEps), p2n = fma(bu, bu, kEps);
    return (p1 - p1n) / (p1 * p1n) + (p2 - p2n) / (p2 * p2n);
}

__global__ void __launch_bounds__(544) anneal_swap_pipe_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;
    const int dns = d * ns;
    double* cur = smem_d;                                   // [k][i]
    double* part = smem_d + dns;                            // [2][32] warp partials, by iteration parity
    double* spec = part + 64;                               // [2][4] special terms {rejected r1, r2, accepted r1, r2}
    Decision* decs = reinterpret_cast<Decision*>(spec + 8); // [2] by iteration parity
    uint4* draws = reinterpret_cast<uint4*>(spec + 12);     // kRing swap draws
    double2* rows = reinterpret_cast<double2*>(draws + kRing);  // [2][d] {x[r1][k], x[r2][k]}, by iteration parity
    ushort4* mlog = reinterpret_cast<ushort4*>(rows + 2 * d);   // kLogCap accepted moves {r1, r2, c}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5, wwarps = nwarps - 1, workers = wwarps * 32;
    const bool decider = warp == wwarps;
    const int npairs = (n + 1) / 2;

    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        if (ns != n)
            for (int k = threadIdx.x; k < d; k += blockDim.x) cur[k * ns + n] = 0.0;
        int ok = 1;
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            const int i = e / d, k = e - i * d;
            const double v = p.design[e];
            ok &= (v >= 0.0 && v <= 1.0) ? 1 : 0;
            cur[k * ns + i] = v;  // anneal.rs:89
        }
        fill_draws(draws, 0ull, kRing, blockDim.x, key0, key1);
        const bool in_cube = __syncthreads_and(ok) != 0;
        publish_partial<true>(full_score_partial<true, 4>(cur, n, d, ns, in_cube), part);
        __syncthreads();
        // chain state, kept by the decider warp (every lane holds the same values)
        double raw_cur = 0.0, cur_metric = 0.0, global_best = 0.0;
        double temp = p.t0;                                                             // anneal.rs:81
        bool materialized = false, overflow = false;
        int log_n = 0, prev_applied = 0;
        if (decider) {
            raw_cur = warp_total<true>(part, nwarps);
            cur_metric = metric_from_raw<true>(raw_cur, n, p.n_pairs, p.inv_d);         // :90
            global_best = cur_metric;                                                   // :94
        }
        __syncthreads();
        double* best_out = p.best_design + chain * (unsigned long long)nd;
        unsigned long long until_resync = p.resync;
        bool have_bulk = false, use_fix = false;

        for (unsigned long long s = 0; s < p.iters; ++s) {
            const int cb = (int)(s & 1), nb = cb ^ 1;
            const Move m = load_move(draws, s, n, d);
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
                } else if (!m.noop && !decider) {
                    for (int k = threadIdx.x; k < d; k += workers) rows[cb * d + k] = make_double2(cur[k * ns + m.r1], cur[k * ns + m.r2]);
                    workers_sync(workers);
                    publish_partial<true>(swap_delta_partial(cur, rows + cb * d, n, d, ns, m.r1, m.r2, m.c, in_cube, -1, -1, workers),
                                          part + 32 * cb);
                }
                __syncthreads();
            }
            // ---- may the workers run ahead?  Not past a re-sum, not when the next move shares a row
            Move mn = m;
            bool next_ok = false;
            if (s + 1 < p.iters) {
                mn = load_move(draws, s + 1, n, d);
                const bool shares_row = !m.noop && !mn.noop && (mn.r1 == m.r1 || mn.r1 == m.r2 || mn.r2 == m.r1 || mn.r2 == m.r2);
                next_ok = !full && until_resync != 1 && !shares_row;
            }
            if (!decider) {
                if (((unsigned int)s & (kDrawBatch - 1)) == 0 && s != 0)
                    fill_draws(draws, s + kDrawBatch, kDrawBatch, workers, key0, key1);  // the half nobody reads any more
                if (next_ok && !mn.noop) {
                    double2* rn = rows + nb * d;
                    for (int k = threadIdx.x; k < d; k += workers) rn[k] = make_double2(cur[k * ns + mn.r1], cur[k * ns + mn.r2]);
                    workers_sync(workers);
                    const int e1 = m.noop ? -1 : m.r1, e2 = m.noop ? -1 : m.r2;
                    publish_partial<true>(swap_delta_partial(cur, rn, n, d, ns, mn.r1, mn.r2, mn.c, in_cube, e1, e2, workers),
                                          part + 32 * nb);
                    if (!m.noop) {
                        // the two rows of move s, under both outcomes: threads past the last pair if there are any
                        const int q = (int)threadIdx.x - (npairs + 4 <= workers ? npairs : 0);
                        if (q >= 0 && q < 4)
                            spec[nb * 4 + q] = special_term(rows + cb * d, rn, d, m.c, mn.c, q & 1, q >> 1);
                    }
                }
            } else {
                // ---- Metropolis (anneal.rs:110-128) and the global best (:131-136)
                double raw_new;
                if (full) {
                    raw_new = warp_total<true>(part + 32 * cb, nwarps);
                } else {
                    double delta = 0.0;
                    if (!m.noop) {
                        delta = warp_total<true>(part + 32 * cb, wwarps);
                        if (use_fix) delta += spec[cb * 4 + 2 * prev_applied] + spec[cb * 4 + 2 * prev_applied + 1];
                    }
                    raw_new = raw_cur + delta;
                }
                const double new_metric = metric_from_raw<true>(raw_new, n, p.n_pairs, p.inv_d);
                double diff = new_metric - cur_metric;
                if (!p.minimize) diff *= -1.0;
                const double p_acc = diff < 0.0 ? 1.0 : exp(-diff / temp);
                const bool accept = m.u < p_acc;
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
        if (decider && lane == 0) {
            p.best_metric[chain] = global_best;
            p.improved[chain] = materialized ? 1u : 0u;
        }
    }
}

size_t pipe_smem_bytes(unsigned long long n, unsigned long long d) {
    const unsigned long long ns = (n + 1) & ~1ull;
    return (size_t)(ns * d + 64 + 8 + 4 + 4 * d) * sizeof(double) + (size_t)kRing * sizeof(uint4) + (size_t)kLogCap * sizeof(ushort4);
}

}  // namespace

bool anneal_pipe_supported(unsigned long long n, unsigned long long d, int metric, int swap) {
    if (const char* e = getenv("MAXPRO_NO_ANNEAL_PIPE")) if (atoi(e) != 0) return false;
    return metric == 0 && swap && n >= 4 && n <= 65535 && d >= 1 && d <= 65535 && pipe_smem_bytes(n, d) <= kSmemBudget;
}

cudaError_t launch_anneal_pipe(const AnnealLaunch& L) {
    AnnealParams p;
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.01 / (double)L.n;
    p.minimize = L.minimize; p.swap = 1;
    unsigned long long rs = 16 * L.n;
    p.resync = rs < 1024 ? 1024 : rs;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved;
    const size_t smem = pipe_smem_bytes(L.n, L.d);
    // worker warps: one thread per two partner rows, plus four threads for the special terms
    int wwarps = (int)(((L.n + 1) / 2 + 4 + 31) / 32);
    if (wwarps > 16) wwarps = 16;
    if (const char* e = getenv("MAXPRO_ANNEAL_THREADS")) { int v = atoi(e) / 32; if (v >= 1 && v <= 16) wwarps = v; }
    const int threads = (wwarps + 1) * 32;
    unsigned long long grid = L.n_chains < (1ull << 20) ? L.n_chains : (1ull << 20);
    cudaError_t e = cudaFuncSetAttribute(anneal_swap_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    anneal_swap_pipe_kernel<<<(unsigned int)grid, threads, smem, L.stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

// K3: many independent Metropolis chains, one CTA per chain.
//
// Replaces anneal::anneal_lhd + swap_rows (src/anneal.rs:23-143).  The reference
// runs ONE chain and re-scores the whole design (O(n^2 d)) every iteration; here
//   - each chain keeps its current design in shared memory, dimension-major
//     [k][i] with an even row stride, for the whole run (160 KB at n=1000, d=20:
//     one chain per SM);
//   - swap moves on the MaxPro criterion are scored incrementally: exchanging
//     x[r1][c] and x[r2][c] only changes the 2(n-2) pairs that touch r1 or r2.
//     A thread takes two neighbouring partner rows j, j+1 (one 16-byte load per
//     dimension), carries the four products over the d-1 untouched dimensions in
//     registers, and evaluates each change as (p - p')/(p p') so the update does
//     not cancel; the four changes of a thread share ONE divide.  S is re-summed
//     from scratch every kResync iterations to stop drift;
//   - the swap draws (r1, r2, c, u) of 128 consecutive iterations are produced in
//     one go, one Philox call per thread, and read back as a broadcast;
//   - jitter moves (every coordinate changes) and maximin moves are re-scored in
//     full: work items of (row i, 8 consecutive cyclic shifts), the 8 running
//     products in registers across the dimensions;
//   - the chain's global best (anneal.rs:92-94,131-136) lives in HBM and is
//     touched only when it improves; when the current design was already the
//     best, a swap improvement rewrites two elements instead of the design.
// Control flow per iteration is the reference's: move, score, diff (negated when
// maximising), p = diff < 0 ? 1 : exp(-diff/temp), one uniform draw ALWAYS
// consumed, accept if u < p, strict-improvement global best, temp *= cooling.
// Draws: Philox key = chain stream (seed + chain), ctr = {t_lo, t_hi, block, domain}.
#include "anneal.cuh"

#include <algorithm>

#include <cstdlib>
#include <type_traits>

#ifndef ANN_PROFILE
#define ANN_PROFILE 0  // 1: clock64 sums per phase of a jitter step, printed by thread 0 of chain 0 (make EXTRA=-DANN_PROFILE=1)
#endif
#if ANN_PROFILE
#include <cstdio>
#define ANN_T(i) { const long long now_ = clock64(); tsum[i] += now_ - tprev; tprev = now_; }
#else
#define ANN_T(i)
#endif

namespace mp {

template <bool MAXPRO, int PLACE, bool SWAP>
__global__ void __launch_bounds__(512) anneal_kernel(AnnealParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int n = (int)p.n, d = (int)p.d, nd = n * d;
    const int ns = (n + 1) & ~1;                      // even row stride: 16-byte partner loads
    const int dns = d * ns;
    // Where the chain's design lives (PLACE, chosen by the host from the shared-memory budget):
    //   0  in shared memory: the current design, and for jitter moves the proposal beside it;
    //   1  jitter moves on a design that fits shared memory only once (n*d*8 B > ~113 KB, e.g. n=1000,
    //      d=20): the proposal takes the shared memory, the current design lives in a per-CTA scratch
    //      buffer in HBM, read once per iteration to build the proposal and written once per
    //      ACCEPTED move -- 2 x 160 KB per step beside a full O(n^2 d) re-score;
    //   2  a design that does not fit shared memory at all (n*d*8 B > 227 KB): current design and
    //      proposal both in the scratch buffer, every load served by L2/L1.  Slow, but any size the
    //      reference accepts can be annealed.
    // PLACE is a template parameter so that every pointer keeps a single address space (a run-time
    // choice between shared and global turns the scoring loads into generic ones: 15 % slower).
    double* const scratch = PLACE != 0 ? p.cur_global + (size_t)blockIdx.x * (PLACE == 2 ? 2 : 1) * dns : nullptr;
    double* cur = PLACE != 0 ? scratch : smem_d;   // [k][i]
    double* cand = PLACE == 2 ? (SWAP ? cur : scratch + dns) : (PLACE == 1 ? smem_d : (SWAP ? cur : smem_d + dns));
    double* part = smem_d + (PLACE == 2 ? 0 : ((PLACE == 1 || SWAP) ? dns : 2 * dns));   // 32 doubles: warp partials of a reduction
    Decision* dec = reinterpret_cast<Decision*>(part + 32);
    uint4* draws = reinterpret_cast<uint4*>(part + 34);   // kDrawBatch swap draws
    double2* rows = reinterpret_cast<double2*>(draws + kDrawBatch);  // [d] {x[r1][k], x[r2][k]} of the move
    ushort4* mlog = reinterpret_cast<ushort4*>(rows + d);            // kLogCap accepted moves {r1, r2, c}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    __shared__ CtaBest cta_best;  // per-CTA fold of the chains' global bests (anneal.cuh)
    if (threadIdx.x == 0) cta_best_reset(&cta_best, p);
    for (unsigned long long chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        const unsigned long long stream = p.seed + chain;
        const uint32_t key0 = (uint32_t)stream, key1 = (uint32_t)(stream >> 32);
        __syncthreads();
        if (ns != n)
            for (int k = threadIdx.x; k < d; k += blockDim.x) { cur[k * ns + n] = 0.0; cand[k * ns + n] = 0.0; }
        int ok = 1;
        for (int e = threadIdx.x; e < nd; e += blockDim.x) {
            const int i = e / d, k = e - i * d;
            const double v = p.design[e];
            ok &= (v >= 0.0 && v <= 1.0) ? 1 : 0;
            cur[k * ns + i] = v;  // anneal.rs:89
        }
        // reciprocal chains and shared denominators are range-safe only in the unit cube
        const bool in_cube = __syncthreads_and(ok) != 0;
        // chain state, kept by warp 0 (every lane holds the same values)
        double raw_cur = 0.0, cur_metric = 0.0, global_best = 0.0;
        double temp = p.t0;                                                             // anneal.rs:81
        bool materialized = false, overflow = false;  // HBM best exists / the move log lost entries
        int log_n = 0;                                // moves accepted since the HBM best equalled cur
        if (warp == 0) {
            raw_cur = *p.raw0;  // the start design's raw criterion, computed once for all chains
            cur_metric = metric_from_raw<MAXPRO>(raw_cur, n, p.n_pairs, p.inv_d);       // :90
            global_best = cur_metric;                                                   // :94
        }
        double* best_out = cta_best_slot(p, &cta_best, nd);
        unsigned long long until_resync = p.resync;  // iterations until the next full re-sum of S

#if ANN_PROFILE
        long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = clock64();
#endif
        for (unsigned long long t = 0; t < p.iters; ++t) {
            // ---- propose and score: every thread ends up with its share `mine` of the reduction
            double mine, u = 0.0;
            int r1 = 0, r2 = 0, c = 0;
            bool incremental = false, applied = false;
            if (SWAP) {
                const unsigned int slot = (unsigned int)t & (kDrawBatch - 1);
                if (slot == 0) {
                    // every reader of the previous batch is past a barrier of iteration t-1
                    for (int q = threadIdx.x; q < kDrawBatch; q += blockDim.x) {
                        const unsigned long long tq = t + (unsigned long long)q;
                        const Philox4 w = philox4x32_10((uint32_t)tq, (uint32_t)(tq >> 32), 0u, kDomainSwap, key0, key1);
                        draws[q] = make_uint4(w.x, w.y, w.z, w.w);
                    }
                    __syncthreads();
                }
                const uint4 w = draws[slot];
                r1 = (int)mulhi_bound(w.x, (uint32_t)n);  // anneal.rs:29
                r2 = (int)mulhi_bound(w.y, (uint32_t)n);  // :30
                c = (int)mulhi_bound(w.z, (uint32_t)d);   // :31
                u = (double)w.w * 0x1p-32;
                if (MAXPRO && --until_resync != 0) {
                    // incremental: only pairs (r1,j), (r2,j), j not in {r1,r2}, change
                    incremental = true;
                    mine = 0.0;
                    if (r1 != r2) {
                        // the previous iteration's readers of rows[] are past its barriers
                        for (int k = threadIdx.x; k < d; k += blockDim.x) rows[k] = make_double2(cur[k * ns + r1], cur[k * ns + r2]);
                        __syncthreads();
                        mine = swap_delta_partial(cur, rows, n, d, ns, r1, r2, c, in_cube, -1, -1, blockDim.x);
                    }
                    publish_partial<true>(mine, part);
                } else {
                    until_resync = p.resync;
                    // maximin, or the periodic full re-sum: apply the move, score; warp 0 undoes a rejected one
                    applied = true;
                    __syncthreads();  // the previous iteration's best write-out still reads cur
                    if (threadIdx.x == 0) {
                        double v1 = cur[c * ns + r1], v2 = cur[c * ns + r2];
                        cur[c * ns + r1] = v2; cur[c * ns + r2] = v1;  // anneal.rs:34-39
                    }
                    __syncthreads();
                    publish_partial<MAXPRO>(n >= kTileScoreMinN ? full_score_tiles<MAXPRO>(cur, n, d, ns, in_cube)
                                                                : full_score_partial<MAXPRO>(cur, n, d, ns, in_cube), part);
                }
            } else {
                // jitter: element e = i*d + k (row-major, the reference's iteration order,
                // anneal.rs:102-107) takes word e%4 of block e/4
                const double two_step = __dadd_rn(p.step, p.step);
                int jok = 1;
                // the accept draw first: its Philox call overlaps the proposal's instead of following it
                Philox4 wa{0u, 0u, 0u, 0u};
                if (warp == 0) wa = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 0u, kDomainJacc, key0, key1);
                auto jitter_element = [&](uint32_t word, int i, int k) {
                    const double uu = (double)word * 0x1p-32;
                    const double delta = __dsub_rn(__dmul_rn(uu, two_step), p.step);
                    double v = __dadd_rn(cur[k * ns + i], delta);
                    jok &= (v == v) ? 1 : 0;  // clamp passes NaN through
                    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);  // clamp(0.0, 1.0)
                    cand[k * ns + i] = v;
                };
                if (nd <= (int)blockDim.x) {
                    // A design with no more elements than the block has threads: one element per thread.  The (up
                    // to) four threads that share a Philox block each repeat its call -- the lanes are idle
                    // otherwise, and a thread's walk over four elements one after the other was most of the
                    // proposal's latency on a single small chain (the CLI's case).
                    const int e = threadIdx.x;
                    if (e < nd) {
                        const Philox4 w = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)(e >> 2), kDomainJit, key0, key1);
                        const int q = e & 3;
                        const uint32_t word = q == 0 ? w.x : (q == 1 ? w.y : (q == 2 ? w.z : w.w));
                        const int i = e / d;
                        jitter_element(word, i, e - i * d);
                    }
                } else {
                    for (int blk = threadIdx.x; blk * 4 < nd; blk += blockDim.x) {
                        const Philox4 w = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)blk, kDomainJit, key0, key1);
                        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
                        int i = (blk * 4) / d, k = blk * 4 - i * d;  // one division per block; the elements walk on from there
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (blk * 4 + q < nd) {
                                jitter_element(ws[q], i, k);
                                if (++k == d) { k = 0; ++i; }
                            }
                        }
                    }
                }
                if (warp == 0) u = (double)wa.x * 0x1p-32;
                ANN_T(0)
                // every coordinate of the proposal has just been clamped into [0, 1]
                const bool cand_in_cube = __syncthreads_and(jok) != 0;
                ANN_T(1)
                const double sc = n >= kTileScoreMinN ? full_score_tiles<MAXPRO>(cand, n, d, ns, cand_in_cube)
                                                      : full_score_partial<MAXPRO>(cand, n, d, ns, cand_in_cube);
                ANN_T(2)
                publish_partial<MAXPRO>(sc, part);
            }
            ANN_T(3)
            __syncthreads();  // partials are in; nobody reads cur / cand for scoring any more
            ANN_T(4)

            // ---- Metropolis (anneal.rs:110-128) and the global best (:131-136), warp 0 only
            if (warp == 0) {
                const double total = warp_total<MAXPRO>(part, (blockDim.x + 31) >> 5);
                const double raw_new = incremental ? raw_cur + total : total;
                const double new_metric = metric_from_raw<MAXPRO>(raw_new, n, p.n_pairs, p.inv_d);
                double diff = new_metric - cur_metric;
                if (!p.minimize) diff *= -1.0;
                const double p_acc = diff < 0.0 ? 1.0 : exp(-diff / temp);
                const bool accept = u < p_acc;
                if (SWAP && accept != applied && lane == 0) {
                    // apply an accepted incremental move / undo a rejected applied one
                    double v1 = cur[c * ns + r1], v2 = cur[c * ns + r2];
                    cur[c * ns + r1] = v2; cur[c * ns + r2] = v1;
                }
                if (accept) { cur_metric = new_metric; raw_cur = raw_new; }
                const bool improves = (p.minimize && cur_metric < global_best) || (!p.minimize && cur_metric > global_best);
                int best_mode = 0;
                if (accept && SWAP) {
                    if (log_n < kLogCap) {
                        if (lane == 0) mlog[log_n] = make_ushort4((unsigned short)r1, (unsigned short)r2, (unsigned short)c, 0);
                        ++log_n;
                    } else {
                        overflow = true;
                    }
                }
                if (improves) {
                    best_mode = (SWAP && materialized && !overflow) ? 1 : 2;
                    if (lane == 0) { dec->log_n = log_n; }
                    global_best = cur_metric;
                    materialized = true; overflow = false; log_n = 0;
                }
                if (lane == 0) { dec->accept = accept ? 1 : 0; dec->best_mode = best_mode; }
                temp *= p.cooling;  // anneal.rs:139
            }
            ANN_T(5)
            __syncthreads();
            ANN_T(6)
            const int accept = dec->accept, best_mode = dec->best_mode;
            if (accept && !SWAP) {
                for (int e = threadIdx.x; e < dns; e += blockDim.x) cur[e] = cand[e];
                __syncthreads();
            }
            if (best_mode == 1) {
                // the HBM best differs from cur only where a logged move touched it
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
            ANN_T(7)
        }
        if (threadIdx.x == 0) cta_best_commit(&cta_best, p, chain, global_best, materialized);
#if ANN_PROFILE
        if (!SWAP && (threadIdx.x == 0 || threadIdx.x == 40) && chain == 0)
            printf("tid %d, %llu jitter steps, cycles per step: propose %lld, barrier %lld, score %lld, warp sum+publish %lld, barrier %lld, metropolis %lld, barrier %lld, rest (copy, best) %lld\n",
                   (int)threadIdx.x, p.iters, tsum[0] / (long long)p.iters, tsum[1] / (long long)p.iters, tsum[2] / (long long)p.iters, tsum[3] / (long long)p.iters,
                   tsum[4] / (long long)p.iters, tsum[5] / (long long)p.iters, tsum[6] / (long long)p.iters, tsum[7] / (long long)p.iters);
#endif
    }
}

// The raw criterion of the start design (anneal.rs:90), once per call: one CTA, same layout and
// summation as the in-chain re-sum.
template <bool MAXPRO, bool IN_HBM>
__global__ void __launch_bounds__(512) anneal_raw0_kernel(const double* design, int n, int d, double* raw0, double* scratch) {
    extern __shared__ __align__(16) double smem_d[];
    const int ns = (n + 1) & ~1;
    double* cur = IN_HBM ? scratch : smem_d;
    double* part = IN_HBM ? smem_d : smem_d + d * ns;
    if (ns != n)
        for (int k = threadIdx.x; k < d; k += blockDim.x) cur[k * ns + n] = 0.0;
    int ok = 1;
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) {
        const int i = e / d, k = e - i * d;
        const double v = design[e];
        ok &= (v >= 0.0 && v <= 1.0) ? 1 : 0;
        cur[k * ns + i] = v;
    }
    const bool in_cube = __syncthreads_and(ok) != 0;
    publish_partial<MAXPRO>(n >= kTileScoreMinN ? full_score_tiles<MAXPRO>(cur, n, d, ns, in_cube)
                                                : full_score_partial<MAXPRO>(cur, n, d, ns, in_cube), part);
    __syncthreads();
    if (threadIdx.x < 32) {
        const double raw = warp_total<MAXPRO>(part, (blockDim.x + 31) >> 5);
        if (threadIdx.x == 0) *raw0 = raw;
    }
}

static bool anneal_all_in_hbm(unsigned long long n, unsigned long long d);

cudaError_t launch_anneal_raw0(const AnnealLaunch& L) {
    const unsigned long long ns = (L.n + 1) & ~1ull;
    const bool hbm = anneal_all_in_hbm(L.n, L.d);
    if (hbm && !L.cur_scratch) return cudaErrorInvalidValue;
    const size_t smem = (size_t)((hbm ? 0 : ns * L.d) + 32) * sizeof(double);
    int threads = (int)((L.n + 31) / 32 * 32);
    if (threads > 512) threads = 512;
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kernel<<<1, threads, smem, L.stream>>>(L.design, (int)L.n, (int)L.d, L.raw0, L.cur_scratch);
    };
    if (L.metric == 0) { if (hbm) go(anneal_raw0_kernel<true, true>); else go(anneal_raw0_kernel<true, false>); }
    else { if (hbm) go(anneal_raw0_kernel<false, true>); else go(anneal_raw0_kernel<false, false>); }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

static size_t anneal_smem_bytes_buffers(unsigned long long n, unsigned long long d, int buffers) {
    const unsigned long long ns = (n + 1) & ~1ull;
    return (size_t)(ns * d * buffers + 34 + 2 * d) * sizeof(double) + (size_t)kDrawBatch * sizeof(uint4) + (size_t)kLogCap * sizeof(ushort4);
}

// Placement of a chain's design (see anneal_kernel): 0 shared memory, 1 jitter with the current design in
// HBM scratch, 2 everything in HBM scratch.
static bool anneal_all_in_hbm(unsigned long long n, unsigned long long d) {
    return anneal_smem_bytes_buffers(n, d, 1) > kSmemBudget - kSmemSlack;
}
static int anneal_place(unsigned long long n, unsigned long long d, int swap) {
    if (anneal_all_in_hbm(n, d)) return 2;
    if (!swap && anneal_smem_bytes_buffers(n, d, 2) > kSmemBudget - kSmemSlack) return 1;
    return 0;
}

size_t anneal_smem_bytes(unsigned long long n, unsigned long long d, int swap) {
    const int place = anneal_place(n, d, swap);
    return anneal_smem_bytes_buffers(n, d, place == 2 ? 0 : ((swap || place == 1) ? 1 : 2));
}

constexpr unsigned long long kHbmGrid = 1024;  // CTAs (= scratch buffers) of a launch that keeps designs in HBM

size_t anneal_jitter_scratch_bytes(unsigned long long n, unsigned long long d, int swap, unsigned long long n_chains) {
    const int place = anneal_place(n, d, swap);
    if (place == 0) return 0;
    const unsigned long long ns = (n + 1) & ~1ull;
    const unsigned long long grid = n_chains < kHbmGrid ? n_chains : kHbmGrid;
    return (size_t)(grid * ns * d * (place == 2 ? 2 : 1)) * sizeof(double) + 64;  // + the masked over-read of the last column block
}

// Every design the 16-bit row indices can address is supported: what does not fit shared memory is annealed out of HBM.
bool anneal_supported(unsigned long long n, unsigned long long d, int swap) {
    return n >= 1 && d >= 1 && n < (1ull << 30) / (d ? d : 1) && n <= 65535 && d <= 65535 &&
           anneal_smem_bytes(n, d, swap) <= kSmemBudget - kSmemSlack;  // the staged rows of a move (2 d doubles) stay in shared memory
}

static int round_up_32(unsigned long long v) { return (int)((v + 31) / 32 * 32); }

static cudaError_t launch_anneal_impl(const AnnealLaunch& L, int plan_sms, unsigned int* plan_grid) {
    if (!plan_grid)
        if (cudaError_t e0 = launch_anneal_raw0(L); e0 != cudaSuccess) return e0;
    if (anneal_pipe_supported(L.n, L.d, L.metric, L.swap)) return launch_anneal_pipe(L, plan_sms, plan_grid);
    if (anneal_maximin_supported(L.n, L.d, L.metric, L.swap)) return launch_anneal_maximin(L, plan_sms, plan_grid);
    AnnealParams p;
    p.design = L.design; p.n = L.n; p.d = L.d; p.iters = L.iters; p.seed = L.seed; p.n_chains = L.n_chains;
    p.t0 = L.t0; p.cooling = L.cooling; p.step = 0.01 / (double)L.n;  // anneal.rs:86-87
    p.minimize = L.minimize; p.swap = L.swap;
    unsigned long long rs = 16 * L.n;
    p.resync = rs < 1024 ? 1024 : rs;
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.best_design = L.best_design; p.best_metric = L.best_metric; p.improved = L.improved; p.raw0 = L.raw0;
    p.cta_chain = L.cta_chain; p.cta_buf = L.cta_buf;
    p.rowmin0 = nullptr; p.rowarg0 = nullptr;
    const int place = anneal_place(L.n, L.d, L.swap);
    if (!plan_grid && place != 0 && !L.cur_scratch) return cudaErrorInvalidValue;
    p.cur_global = place != 0 ? L.cur_scratch : nullptr;
    size_t smem = anneal_smem_bytes(L.n, L.d, L.swap);
    int threads;
    if (L.swap && L.metric == 0) {
        // incremental steps: one thread per two partner rows
        threads = round_up_32((L.n + 1) / 2);
    } else {
        // full re-score every iteration.  Designs on broadcast tiles (n >= kTileScoreMinN): one thread per row.
        // Smaller ones: one thread per work item (row, kShiftBlock shifts) of full_score_partial -- a chain of a
        // 50 x 2 design on 64 threads spent 3,000 of its 7,250 cycles per step walking four items per thread.
        threads = round_up_32(L.n);
        if ((int)L.n < kTileScoreMinN) {
            const unsigned long long smax = (L.n - 1) / 2 + ((L.n & 1) == 0 ? 1 : 0);
            threads = round_up_32(L.n * ((smax + kShiftBlock - 1) / kShiftBlock));
        }
    }
    if (threads < 32) threads = 32;
    if (threads > 512) threads = 512;
    if (const char* e = cfg("MAXPRO_ANNEAL_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 512 && v % 32 == 0) threads = v; }
    const unsigned long long cap = place != 0 ? kHbmGrid : (1ull << 20);  // HBM placements: one scratch buffer per CTA
    cudaError_t e = cudaSuccess;
    // Many maximin chains (more chains than SMs, designs in shared memory): the block size that keeps the most threads
    // resident among the CTAs the chains can fill, the smaller block on a tie -- several small CTAs overlap their barriers and
    // Metropolis steps where one large CTA waits: 20,000 jitter chains of (50,3) 1.5e8 -> 2.9e8 chain-steps/s, (300,6)
    // 0.93e7 -> 1.23e7 (tools/anneal_many_chains_jit.py).  A minimum does not depend on the order it is taken in, so the
    // trajectory is the same at every block size; the MaxPro sums are not order-free and keep their block size.
    int sms = plan_sms;
    if (sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); sms = 0; }
    }
    const bool pick_threads = L.metric != 0 && place == 0 && sms > 0 && L.n_chains > (unsigned long long)sms && !cfg("MAXPRO_ANNEAL_THREADS");
    auto go = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        if (pick_threads) {
            const unsigned long long needed = (L.n_chains + sms - 1) / sms;
            long long best_used = -1;
            for (int t = threads; t >= 32; t -= 32) {
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, t, smem) != cudaSuccess) { cudaGetLastError(); continue; }
                const long long used = (long long)std::min<unsigned long long>((unsigned long long)per_sm, needed) * t;
                if (used >= best_used) { best_used = used; threads = t; }
            }
        }
        if (plan_grid) { e = anneal_resident_grid(kernel, threads, smem, plan_sms, L.n_chains, cap, plan_grid); return; }
        if (L.grid == 0 || L.grid > L.n_chains || L.grid > cap) { e = cudaErrorInvalidValue; return; }
        kernel<<<L.grid, threads, smem, L.stream>>>(p);
    };
    auto pick = [&](auto mp_tag) {
        constexpr bool MP = decltype(mp_tag)::value;
        if (L.swap) { if (place == 2) go(anneal_kernel<MP, 2, true>); else go(anneal_kernel<MP, 0, true>); }
        else if (place == 2) go(anneal_kernel<MP, 2, false>);
        else if (place == 1) go(anneal_kernel<MP, 1, false>);
        else go(anneal_kernel<MP, 0, false>);
    };
    if (L.metric == 0) pick(std::true_type{}); else pick(std::false_type{});
    if (e != cudaSuccess || plan_grid) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_anneal(const AnnealLaunch& L) { return launch_anneal_impl(L, 0, nullptr); }

cudaError_t anneal_plan_grid(const AnnealLaunch& L, int sms, unsigned int* grid) { return launch_anneal_impl(L, sms, grid); }

}  // namespace mp

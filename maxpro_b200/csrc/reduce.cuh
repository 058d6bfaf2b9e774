// Block-level argbest + last-CTA fold shared by the search and scoring kernels.
#pragma once
#include "common.cuh"
#include "pow_glibc.cuh"
#include <math_constants.h>

namespace mp {

// psi = (S / n_pairs)^(1/d), src/maxpro_utils.rs:81-82, rounded like the libm pow the reference calls.
__device__ __forceinline__ double maxpro_psi(double s, double n_pairs, double inv_d) { return pow_glibc(s / n_pairs, inv_d); }

// A lane's running best under the fold of build_lhd (src/lib.rs:85-97).  The reference compares the
// CRITERION -- psi after powf for MaxPro -- and keeps the later index when two criteria are equal.  Taking
// the power of every candidate would cost more than scoring it, so a MaxPro lane keeps, beside its best
// psi, a bound `thr` on the raw sum S above which psi is certainly larger: psi(S (1 + t)) ~ psi(S) (1 + t/d),
// so t = (4 d + 4) 2^-52 puts psi(S (1 + t)) at least two ulp above psi(S), beyond anything the rounding of
// pow (< 1 ulp, monotone to that accuracy) can undo.  Only a candidate with S <= thr -- a new best of the
// lane or a near-tie with it, a handful per lane and launch -- pays for pow_glibc and is compared on psi
// exactly as the reference does.  Maximin criteria are compared as they are (sqrt of the minimum: exact ties).
template <bool MAXPRO>
struct LaneBest {
    double best, thr, tie;
    unsigned long long idx;
    __device__ __forceinline__ void init(double inv_d) {
        best = MAXPRO ? CUDART_INF : -CUDART_INF;
        thr = CUDART_INF;
        tie = fma(4.0 / inv_d + 4.0, 0x1p-52, 1.0);
        idx = kNoIndex;
    }
    // MaxPro: s = raw sum S.  Maximin: s = the criterion itself.
    __device__ __forceinline__ void offer(double s, unsigned long long i, double n_pairs, double inv_d) {
        if (MAXPRO) {
            if (s <= thr) {  // false for NaN, as `m1 < m2` is in the reference
                const double psi = maxpro_psi(s, n_pairs, inv_d);
                if (psi < best || (psi == best && (idx == kNoIndex || i > idx))) { best = psi; idx = i; thr = s * tie; }
            }
        } else {
            offer_criterion(s, i);
        }
    }
    // c = the criterion itself (psi or the maximin distance); smaller wins for MaxPro, larger for maximin
    __device__ __forceinline__ void offer_criterion(double c, unsigned long long i) {
        if ((MAXPRO ? c < best : c > best) || (c == best && (idx == kNoIndex || i > idx))) { best = c; idx = i; }
    }
};

// Block-level argbest, then the last CTA to finish folds the per-CTA records:
// only 16 bytes per CTA ever reach HBM.
template <bool MINIMIZE>
__device__ __forceinline__ void block_publish_best(double s, unsigned long long idx, BestRec* warp_best,
                                                   BestRec* block_best, unsigned int* done_counter,
                                                   BestRec* result) {
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    warp_reduce_best<MINIMIZE>(s, idx);
    if (lane == 0) { warp_best[warp].score = s; warp_best[warp].index = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarps; ++w)
            if (better<MINIMIZE>(warp_best[w].score, warp_best[w].index, s, idx)) {
                s = warp_best[w].score; idx = warp_best[w].index;
            }
        block_best[blockIdx.x].score = s;
        block_best[blockIdx.x].index = idx;
        __threadfence();
        unsigned int prev = atomicAdd(done_counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double bs = MINIMIZE ? CUDART_INF : -CUDART_INF;
    unsigned long long bi = kNoIndex;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const volatile BestRec* rec = block_best + b;
        double so = rec->score; unsigned long long io = rec->index;
        if (better<MINIMIZE>(so, io, bs, bi)) { bs = so; bi = io; }
    }
    warp_reduce_best<MINIMIZE>(bs, bi);
    __syncthreads();
    if (lane == 0) { warp_best[warp].score = bs; warp_best[warp].index = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarps; ++w)
            if (better<MINIMIZE>(warp_best[w].score, warp_best[w].index, bs, bi)) {
                bs = warp_best[w].score; bi = warp_best[w].index;
            }
        result->score = bs;
        result->index = bi;
        *done_counter = 0u;  // re-arm for the next launch on this workspace
    }
}


}  // namespace mp

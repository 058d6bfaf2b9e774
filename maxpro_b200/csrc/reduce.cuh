// Block-level argbest + last-CTA fold shared by the search and scoring kernels.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace mp {

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

// Scoring of a design held dimension-major ([k][i], even row stride) in work items of 64 rows x 8
// columns: lane <-> rows row0+lane and row0+lane+32, the 8 column coordinates are 16-byte warp-wide
// broadcasts, the 16 running products (or squared distances) stay in registers across the
// dimensions.  8 shared-memory wavefronts per dimension for 16 pairs per lane -- the cheapest of
// the engine's schedules in shared-memory bandwidth, which is what the annealer's full re-score
// (one row x 8 cyclic shifts per item: 18 wavefronts for 8 pairs) is bound by.
// Used by the CTA-per-candidate search (search_cta.cu) and by the annealer's full re-score of
// designs with n >= 96 (anneal.cuh).
#pragma once
#include "common.cuh"

namespace mp {

constexpr int kTileRows = 64;  // 2 rows per lane
constexpr int kJB = 8;         // columns per block

// Eight consecutive partner coordinates as four 16-byte warp-wide broadcasts.  The row stride is
// even and a block starts at a multiple of 8, so the address is 16-byte aligned; a block that
// runs past row n-1 reads the stride padding (or the next dimension) and is masked by the caller.
__device__ __forceinline__ void load_block(double (&bv)[kJB], const double* src) {
#pragma unroll
    for (int c = 0; c < kJB; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + c);
        bv[c] = v.x; bv[c + 1] = v.y;
    }
}

// Per-lane scoring state of one candidate: eight reciprocal chains (one per block column), the
// number of terms each holds, the running sum (MaxPro) or the running minimum (maximin).
struct CtaScore {
    RecipChain ch[kJB];
    int pushed;
    double acc, mn;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int c = 0; c < kJB; ++c) ch[c].reset();
        pushed = 0; acc = 0.0; mn = CUDART_INF;
    }
    __device__ __forceinline__ void flush() {
#pragma unroll
        for (int c = 0; c < kJB; ++c) acc = ch[c].flush_into(acc);
        pushed = 0;
    }
};

// One dimension of a work item: ROWS x 8 differences per lane (one or two rows against eight broadcast
// columns), multiplied into (MaxPro) or squared and added to (maximin) the running values.
template <bool MAXPRO, bool FIRST, int ROWS>
__device__ __forceinline__ void cta_item_dim(double (&prod)[2][kJB], const double* xk, int i0c, int i1c, int j0) {
    double a[2];
    a[0] = xk[i0c];
    if (ROWS == 2) a[1] = xk[i1c];
    double bv[kJB];
    load_block(bv, xk + j0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < kJB; ++c) {
            if (MAXPRO) {
                if (FIRST) prod[r][c] = a[r] - bv[c];
                else prod[r][c] *= (a[r] - bv[c]);
            } else {
                const double df = __dsub_rn(a[r], bv[c]);
                if (FIRST) prod[r][c] = __dmul_rn(df, df);
                else prod[r][c] = __dadd_rn(prod[r][c], __dmul_rn(df, df));
            }
        }
}

// One work item: the 64 rows of tile row0 (lane <-> rows row0+lane, row0+lane+32) against the 8
// columns j0 .. j0+7.  The dimensions are walked in fully unrolled chunks of U (the caller picks a
// U that divides d when one of 5, 4, 3 does, so that d = 10 or 20 runs two or four identical
// chunks with no peeled first dimension and no remainder loop); what is left goes one at a time.
// ROWS = 1 serves the diagonal blocks that lie in the first half of their own tile: every row of
// the second half follows all eight columns there, so only the lane's first row is evaluated.
template <bool MAXPRO, int U, int ROWS, bool SAFE = false>
__device__ __forceinline__ void cta_score_item(const double* xs, int n, int d, int ns, int row0, int j0, int lane, CtaScore& st) {
    const int i0 = row0 + lane, i1 = i0 + 32;
    const int i0c = min(i0, n - 1), i1c = min(i1, n - 1);
    double prod[2][kJB];
    int k = 0;
    if (d >= U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u == 0) cta_item_dim<MAXPRO, true, ROWS>(prod, xs, i0c, i1c, j0);
            else cta_item_dim<MAXPRO, false, ROWS>(prod, xs + u * ns, i0c, i1c, j0);
        }
        k = U;
#pragma unroll 1
        for (; k + U <= d; k += U) {
            const double* xk = xs + k * ns;
#pragma unroll
            for (int u = 0; u < U; ++u) cta_item_dim<MAXPRO, false, ROWS>(prod, xk + u * ns, i0c, i1c, j0);
        }
    } else {
        cta_item_dim<MAXPRO, true, ROWS>(prod, xs, i0c, i1c, j0);
        k = 1;
    }
#pragma unroll 1
    for (; k < d; ++k) cta_item_dim<MAXPRO, false, ROWS>(prod, xs + k * ns, i0c, i1c, j0);
    // full blocks: every row of the tile precedes every column and all are < n
    const bool full = ROWS == 2 && (j0 >= row0 + kTileRows) && (j0 + kJB <= n) && (row0 + kTileRows <= n);
    if (MAXPRO && SAFE) {
        // designs outside the unit cube: the reciprocal chains are not range-safe, every term takes a divide
#pragma unroll
        for (int c = 0; c < kJB; ++c) {
            const int j = j0 + c;
            if (full || (j < n && i0 < j)) st.acc += 1.0 / fma(prod[0][c], prod[0][c], kEps);
            if (ROWS == 2 && (full || (j < n && i1 < j))) st.acc += 1.0 / fma(prod[1][c], prod[1][c], kEps);
        }
    } else if (MAXPRO) {
        if (st.pushed + ROWS > kMaxChainTerms) st.flush();
        if (full) {
#pragma unroll
            for (int c = 0; c < kJB; ++c) {
                st.ch[c].push(fma(prod[0][c], prod[0][c], kEps));
                st.ch[c].push(fma(prod[1][c], prod[1][c], kEps));
            }
        } else {
#pragma unroll
            for (int c = 0; c < kJB; ++c) {
                const int j = j0 + c;
                if (j < n && i0 < j) st.ch[c].push(fma(prod[0][c], prod[0][c], kEps));
                if (ROWS == 2 && j < n && i1 < j) st.ch[c].push(fma(prod[1][c], prod[1][c], kEps));
            }
        }
        st.pushed += ROWS;
    } else {
#pragma unroll
        for (int c = 0; c < kJB; ++c) {
            const int j = j0 + c;
            if (full || (j < n && i0 < j)) st.mn = min_sqdist(st.mn, prod[0][c]);
            if (ROWS == 2 && (full || (j < n && i1 < j))) st.mn = min_sqdist(st.mn, prod[1][c]);
        }
    }
}

// chunk size for cta_score_item: the largest of 5, 4, 3 that divides d, else 4
__host__ __device__ __forceinline__ int cta_dim_chunk(int d) { return d % 5 == 0 ? 5 : d % 4 == 0 ? 4 : d % 3 == 0 ? 3 : 4; }

template <bool MAXPRO, bool SAFE = false>
__device__ __forceinline__ void cta_score_item_any(int chunk, const double* xs, int n, int d, int ns, int row0, int j0, int lane, CtaScore& st) {
    if (j0 + kJB <= row0 + 32) {  // no column of the block follows a row of the tile's second half
        if (chunk == 5) cta_score_item<MAXPRO, 5, 1, SAFE>(xs, n, d, ns, row0, j0, lane, st);
        else if (chunk == 3) cta_score_item<MAXPRO, 3, 1, SAFE>(xs, n, d, ns, row0, j0, lane, st);
        else cta_score_item<MAXPRO, 4, 1, SAFE>(xs, n, d, ns, row0, j0, lane, st);
        return;
    }
    if (chunk == 5) cta_score_item<MAXPRO, 5, 2, SAFE>(xs, n, d, ns, row0, j0, lane, st);
    else if (chunk == 3) cta_score_item<MAXPRO, 3, 2, SAFE>(xs, n, d, ns, row0, j0, lane, st);
    else cta_score_item<MAXPRO, 4, 2, SAFE>(xs, n, d, ns, row0, j0, lane, st);
}

}  // namespace mp

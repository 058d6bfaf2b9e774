// K2 on caller-supplied designs: the parity entry point and the engine behind
// maxpro_criterion / maximin_criterion / score_batch for ANY (n, d).
//
// Replaces maxpro_utils::maxpro_sum + maxpro_criterion (src/maxpro_utils.rs:28-83),
// maximin_utils::maximin_criterion (src/maximin_utils.rs:54-67) and the
// map/reduce of src/lib.rs:65-98 over a batch.
//
// Decomposition: the i<j triangle of one design is cut into 64x64 tile pairs
// (I <= J).  A work item is (design, I, J); a CTA of 256 threads owns an item,
// thread (ty,tx) a 4x4 block of pairs (rows ty+16r, cols tx+16c) so the
// shared-memory reads are broadcasts (rows) or 16 consecutive doubles (cols).
// Tiles are staged dimension-major [k][i] in chunks of kKC dimensions, products
// (or squared distances) stay in registers across chunks.  Per-item partials go
// to HBM (8 B per item) and a second kernel folds them in a fixed order, applies
// (S/pairs)^(1/d) or sqrt, and does the argbest with the reference tie rule, so
// the result does not depend on scheduling.
//
// Arithmetic: maxpro terms use an IEEE divide when the designs come from the caller
// (they may lie outside the unit cube, where the reciprocal chains of the search
// kernels are not range-safe); when the search itself generated them (chunked
// path for designs above 227 KB) the launch says so and the 16 terms of a thread
// go through four reciprocal chains: 3 FP64 instructions per term instead of
// ~12.  Maximin uses separate multiply and add in the reference's k order and
// is bit-exact with src/maximin_utils.rs:39-44.
#include "common.cuh"
#include "reduce.cuh"
#include "kernels.h"

namespace mp {

constexpr int kTS = 64;   // tile side
constexpr int kKC = 16;   // dimensions staged per chunk
constexpr int kLd = kTS + 1;

struct PairParams {
    const double* designs;
    unsigned long long b, n, d;
    unsigned int tiles, items;  // per design
    double* partial;            // [b * items]
};

template <bool MAXPRO, bool IN_CUBE>
__global__ void __launch_bounds__(256) pair_tile_kernel(PairParams p) {
    __shared__ double xi_s[kKC][kLd];
    __shared__ double xj_s[kKC][kLd];
    __shared__ double warp_part[8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const unsigned long long work = p.b * (unsigned long long)p.items;

    for (unsigned long long wi = blockIdx.x; wi < work; wi += gridDim.x) {
        const unsigned long long design = wi / p.items;
        unsigned int rem = (unsigned int)(wi % p.items), I = 0;
        while (rem >= p.tiles - I) { rem -= p.tiles - I; ++I; }
        const unsigned int J = I + rem;
        const double* x = p.designs + design * p.n * p.d;
        const unsigned long long row0 = (unsigned long long)I * kTS, col0 = (unsigned long long)J * kTS;
        const unsigned long long rows = min((unsigned long long)kTS, p.n - row0);
        const unsigned long long cols = min((unsigned long long)kTS, p.n - col0);

        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = MAXPRO ? 1.0 : 0.0;

        for (unsigned long long k0 = 0; k0 < p.d; k0 += kKC) {
            const int kc = (int)min((unsigned long long)kKC, p.d - k0);
            __syncthreads();
            // stage [kc][64] of both tiles; rows beyond n are zero-filled (masked later)
            for (int e = tid; e < kTS * kc; e += 256) {
                int i = e / kc, k = e - i * kc;
                xi_s[k][i] = (unsigned long long)i < rows ? x[(row0 + i) * p.d + k0 + k] : 0.0;
                xj_s[k][i] = (unsigned long long)i < cols ? x[(col0 + i) * p.d + k0 + k] : 0.0;
            }
            __syncthreads();
            for (int k = 0; k < kc; ++k) {
                double a[4], bb[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = xi_s[k][ty + 16 * r];
#pragma unroll
                for (int c = 0; c < 4; ++c) bb[c] = xj_s[k][tx + 16 * c];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        double df = __dsub_rn(a[r], bb[c]);
                        if (MAXPRO) acc[r][c] *= df;  // product of differences, squared once at the end
                        else acc[r][c] = (k0 + k == 0) ? __dmul_rn(df, df) : __dadd_rn(acc[r][c], __dmul_rn(df, df));
                    }
            }
        }

        double t = MAXPRO ? 0.0 : CUDART_INF;
        RecipChain ch[4];
        if (MAXPRO && IN_CUBE) {
#pragma unroll
            for (int r = 0; r < 4; ++r) ch[r].reset();
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const unsigned long long gi = row0 + ty + 16 * r, gj = col0 + tx + 16 * c;
                const bool valid = gi < p.n && gj < p.n && gi < gj;
                if (MAXPRO) {
                    double q = acc[r][c];
                    double pr = fma(q, q, kEps);  // src/maxpro_utils.rs:50-53
                    if (IN_CUBE) { if (valid) ch[r].push(pr); }
                    else if (valid) t += 1.0 / pr;  // :54
                } else if (valid) {
                    t = min_sqdist(t, acc[r][c]);
                }
            }
        if (MAXPRO && IN_CUBE) {
#pragma unroll
            for (int r = 0; r < 4; ++r) t = ch[r].flush_into(t);  // empty chains add 0
        }
        t = MAXPRO ? warp_reduce_sum(t) : warp_reduce_min(t);
        if ((tid & 31) == 0) warp_part[tid >> 5] = t;
        __syncthreads();
        if (tid == 0) {
            double v = warp_part[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) v = MAXPRO ? v + warp_part[w] : fmin(v, warp_part[w]);
            p.partial[wi] = v;
        }
    }
}

struct FinalizeParams {
    const double* partial;
    unsigned long long b, n;
    unsigned int items;
    double n_pairs, inv_d;
    double* scores;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

template <bool MAXPRO>
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeParams p) {
    __shared__ BestRec warp_best[8];
    double best = MAXPRO ? CUDART_INF : -CUDART_INF;
    unsigned long long best_idx = kNoIndex;
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < p.b;
         c += (unsigned long long)gridDim.x * blockDim.x) {
        const double* part = p.partial + c * p.items;
        double s;
        if (MAXPRO) {
            double sum = 0.0;
            for (unsigned int it = 0; it < p.items; ++it) sum += part[it];
            // n < 2 -> 0.0 (src/maxpro_utils.rs:75-77); else (S/pairs)^(1/d) (:81-82)
            s = p.n < 2 ? 0.0 : maxpro_psi(sum, p.n_pairs, p.inv_d);
        } else {
            double m = CUDART_INF;  // n == 1 -> +INF (src/maximin_utils.rs:57)
            for (unsigned int it = 0; it < p.items; ++it) m = fmin(m, part[it]);
            s = sqrt(m);
        }
        p.scores[c] = s;
        if (MAXPRO ? (s <= best) : (s >= best)) { best = s; best_idx = c; }  // later index on ties
    }
    if (p.result)
        block_publish_best<MAXPRO>(best, best_idx, warp_best, p.block_best, p.done_counter, p.result);
}

unsigned long long score_items(unsigned long long n) {
    unsigned long long tiles = (n + kTS - 1) / kTS;
    if (tiles == 0) tiles = 1;
    return tiles * (tiles + 1) / 2;
}

int score_finalize_grid(unsigned long long b) {
    unsigned long long g = (b + 255) / 256;
    if (g > 148 * 4) g = 148 * 4;
    if (g == 0) g = 1;
    return (int)g;
}

cudaError_t launch_score_batch(const ScoreLaunch& L) {
    PairParams p;
    p.designs = L.designs; p.b = L.b; p.n = L.n; p.d = L.d;
    p.tiles = (unsigned int)((L.n + kTS - 1) / kTS);
    if (p.tiles == 0) p.tiles = 1;
    p.items = (unsigned int)score_items(L.n);
    p.partial = L.partial;
    unsigned long long work = L.b * (unsigned long long)p.items;
    int grid = (int)(work < 148ull * 16 ? work : 148ull * 16);
    if (grid == 0) return cudaSuccess;
    if (L.metric == 0) {
        if (L.in_cube) pair_tile_kernel<true, true><<<grid, 256, 0, L.stream>>>(p);
        else pair_tile_kernel<true, false><<<grid, 256, 0, L.stream>>>(p);
    } else {
        pair_tile_kernel<false, false><<<grid, 256, 0, L.stream>>>(p);
    }
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    FinalizeParams f;
    f.partial = L.partial; f.b = L.b; f.n = L.n; f.items = p.items;
    f.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    f.inv_d = 1.0 / (double)L.d;
    f.scores = L.scores; f.block_best = L.block_best; f.done_counter = L.done_counter; f.result = L.result;
    int fgrid = score_finalize_grid(L.b);
    if (L.metric == 0) finalize_kernel<true><<<fgrid, 256, 0, L.stream>>>(f);
    else finalize_kernel<false><<<fgrid, 256, 0, L.stream>>>(f);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

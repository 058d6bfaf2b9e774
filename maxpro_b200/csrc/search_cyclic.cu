// K1+K2 fused, shape-specialised variant of search_small.cu for the headline shapes
// (n = 50, d = 2 or 3): same thread-group-per-candidate layout, but the pair schedule is
// CYCLIC and every loop bound and shared-memory stride is a compile-time constant.
//
// Replaces the same reference code as search_small.cu: generate_lhd (src/lhd.rs:97-132) ->
// criterion (src/maxpro_utils.rs:28-83 / src/maximin_utils.rs:54-67) -> fold (src/lib.rs:76-98).
//
// Why a second kernel.  ncu on the triangular schedule of search_small.cu (profiles/) showed
// one warp alone reaching only ~45% of the FP64 pipe: 18% of the pairs live in remainder
// loops, tile prologues and the in-tile triangle, and those took 42% of the time.  The
// cyclic schedule has ONE block shape and nothing else:
//     pair {i, (i+s) mod n},  s = 1 .. (n-1)/2,  every row i          (+ the n/2 diameters)
// The two lanes of a candidate split the s range in halves (s = 1..12 and 13..24 at n = 50),
// walk the rows in register tiles of R = 5 and read the 5 + 12 - 1 = 16 partner rows of a tile
// once: 60 pairs per tile, 63 LDS.64, fully unrolled, no triangle, no remainder, identical trip
// counts in both lanes (no divergence).  With N and C constant every shared-memory address of
// the lower tiles is base + immediate; the upper tiles pay one add and one unsigned min per
// partner row for the wrap-around.
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cstdlib>
#include <cstdio>

// Compile-time experiment switches (off in the product; profiles/README.md, round 2, has the measurements):
//   CYC_EXP   1 no generation after the first round, 3 generation only, 7 / 8 two scoring warps per sub-partition with
//             the third idle / generating three designs per round;  CYC_KNOCK  bit mask of generator parts knocked out;
//   CYC_PROF  clock64 sums of the generation and scoring phases per warp, printed by CTA 0.
// The variants that tried to OVERLAP the two phases (generation steps issued inside the scoring loop, a generation mutex per
// sub-partition, a staggered start) were measured and removed again: commit 408d3ad has them.
#ifndef CYC_EXP
#define CYC_EXP 0
#endif
#ifndef CYC_KNOCK
#define CYC_KNOCK 0
#endif

namespace mp {

struct CyclicParams {
    int gen;   // design stream version (philox.cuh)
    int ring;  // warps per sub-partition that take turns on the FP64 pipe (0 = no turn-taking)
    unsigned long long seed, lo, hi;
    // list mode (stage 2 of the fp32-screened search): score list[0..list_count) instead of lo..hi
    const unsigned long long* list;
    unsigned long long list_count;
    const unsigned int* list_count_dev;  // list length written by stage 1; above list_cap = overflow: score lo..hi instead
    unsigned int list_cap;
    // supplied mode (score_batch): candidate idx is the design designs[idx*N*D ...], row-major
    // [i][k] in HBM; scores[idx] receives its criterion.  lo..hi index the batch.
    const double* designs;
    double* scores;
    double c1, c0;
    double n_pairs, inv_d;
    BestRec* block_best;
    unsigned int* done_counter;
    BestRec* result;
};

template <int D>
__device__ __forceinline__ double cyc_maxpro_term(const double (&a)[D], const double (&b)[D]) {
    double q = a[0] - b[0];
#pragma unroll
    for (int k = 1; k < D; ++k) q *= (a[k] - b[k]);
    return fma(q, q, kEps);
}

template <int D>
__device__ __forceinline__ double cyc_sqdist_term(const double (&a)[D], const double (&b)[D]) {
    double df = __dsub_rn(a[0], b[0]);
    double s = __dmul_rn(df, df);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        df = __dsub_rn(a[k], b[k]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return s;
}

// Byte offsets inside one candidate's slice: element (k, i) at (k*N + i) * C * 8.
template <int N, int D, int C>
struct Layout {
    static constexpr unsigned kRowBytes = C * 8u;        // i -> i+1
    static constexpr unsigned kDimBytes = N * C * 8u;    // k -> k+1
};

template <int N, int D, int C>
__device__ __forceinline__ void load_row_at(double (&dst)[D], const unsigned char* cand, unsigned row_off) {
#pragma unroll
    for (int k = 0; k < D; ++k)
        dst[k] = *reinterpret_cast<const double*>(cand + row_off + k * Layout<N, D, C>::kDimBytes);
}

// One Fisher-Yates step of the inside-out shuffle: a[i] = a[j]; a[j] = v.
template <int N, int D, int C>
__device__ __forceinline__ void cyc_fy_step(unsigned char* col, int i, uint32_t j, double v) {
    constexpr unsigned RB = Layout<N, D, C>::kRowBytes;
    *reinterpret_cast<double*>(col + i * RB) = *reinterpret_cast<double*>(col + j * RB);
    *reinterpret_cast<double*>(col + j * RB) = v;
}

// Draw blocks [b0, b0 + count) of one column: two Fisher-Yates steps per block.  `col`, `k`
// and `b0` may differ between the two lanes of a candidate; `count` is uniform.
// GEN = 1: LHD-Philox v1, one Philox call per block.  GEN = 2: LHD-mix v2 (philox.cuh), one Philox call per
// column segment for the key, then a multiply-xorshift mix of a Weyl sequence per element.
template <int N, int D, int C, int GEN>
__device__ __forceinline__ void cyc_generate_blocks(unsigned char* col, uint32_t k, int b0, int count, uint32_t s_lo,
                                                    uint32_t s_hi, double c1, double c0) {
    if (GEN == 1) {
#pragma unroll 2
        for (int t = 0; t < count; ++t) {
            const int b = b0 + t;
            const Philox4 w = philox4x32_10((uint32_t)b, k, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
            const int i = 2 * b;
            if (i < N) cyc_fy_step<N, D, C>(col, i, __umulhi(w.y, (uint32_t)i + 1u), stratum_value((uint32_t)i, w.x, c1, c0));
            if (i + 1 < N) cyc_fy_step<N, D, C>(col, i + 1, __umulhi(w.w, (uint32_t)i + 2u), stratum_value((uint32_t)i + 1u, w.z, c1, c0));
        }
    } else {
        uint32_t ka, kb;
        lhd2_column_key(k, s_lo, s_hi, ka, kb);
        uint32_t x = ka + (uint32_t)(2 * b0) * kb;  // Weyl sequence A + i B, stepped instead of multiplied
#pragma unroll 2
        for (int t = 0; t < count; ++t) {
            const int i = 2 * (b0 + t);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                uint32_t j, jit;
                mulhilo32(lhd2_mix32(x), (uint32_t)(i + e) + 1u, j, jit);
                x += kb;
                if (i + e < N) cyc_fy_step<N, D, C>(col, i + e, j, lhd2_value((uint32_t)(i + e), jit, c1, c0));
            }
        }
    }
}

// Three columns on two lanes without an idle lane.  A column's Fisher-Yates chain is sequential,
// but nothing says one lane must walk all of it: lane 0 starts the third column (blocks
// 0..h-1) and then does column 0; lane 1 does column 1 and then finishes the third column
// (blocks h..) -- by then lane 0 has long left it (a __syncwarp orders the hand-over).
// 38 block iterations per lane instead of 50.
template <int N, int C, int GEN>
__device__ __forceinline__ void cyc_generate_three(unsigned char* cand, uint32_t s_lo, uint32_t s_hi, double c1, double c0,
                                                   int g, unsigned int group_mask) {
    constexpr int D = 3;
    constexpr int NB = (N + 1) / 2, H = (NB + 1) / 2;  // blocks per column, blocks of column 2 done by lane 0
    constexpr unsigned DB = Layout<N, D, C>::kDimBytes;
    static_assert(NB - H <= NB && H <= NB, "hand-over happens after lane 0 is done with column 2");
    // stage 1 (H iterations): lane 0 -> column 2 blocks [0,H); lane 1 -> column 1 blocks [0,H)
    cyc_generate_blocks<N, D, C, GEN>(cand + (g ? 1u : 2u) * DB, g ? 1u : 2u, 0, H, s_lo, s_hi, c1, c0);
    // stage 2 (NB-H iterations): lane 0 -> column 0 blocks [0,NB-H); lane 1 -> column 1 blocks [H,NB)
    cyc_generate_blocks<N, D, C, GEN>(cand + (g ? 1u : 0u) * DB, g ? 1u : 0u, g ? H : 0, NB - H, s_lo, s_hi, c1, c0);
    __syncwarp(group_mask);  // column 2's first half is in place before lane 1 continues it
    // stage 3 (H iterations): lane 0 -> column 0 blocks [NB-H,NB); lane 1 -> column 2 blocks [H,NB) (+ one idle)
    cyc_generate_blocks<N, D, C, GEN>(cand + (g ? 2u : 0u) * DB, g ? 2u : 0u, g ? H : NB - H, H, s_lo, s_hi, c1, c0);
}

// ---- LHD-mix v2 generation for the two-lane layout: short dependent chains ---------------------------------------
// A Fisher-Yates step reads a position an EARLIER step of the same column may have written, so a column is a chain
// of dependent shared-memory round trips (~35 cycles each): with Philox gone that latency, not the instruction
// count, is what generation costs.  Two things shorten the chain:
//   * the two elements of a draw block are loaded TOGETHER, before either is stored; what the second one would
//     have read after the first one's stores is patched in from registers (it can only alias a[2b] or a[j0]);
//   * two columns are walked side by side by each lane (lane g: column g all the way, and half of the third
//     column -- lane 0 its first half while t < H, lane 1 the rest afterwards), predicated, not branched.
// 25 block steps per candidate instead of 76 element steps.
__device__ __forceinline__ double cyc_lds(unsigned int addr, bool on, double keep) {
    double t = keep;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.shared.f64 %0, [%1]; }" : "+d"(t) : "r"(addr), "r"((unsigned int)on) : "memory");
    return t;
}
__device__ __forceinline__ void cyc_sts(unsigned int addr, double v, bool on) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(addr), "d"(v), "r"((unsigned int)on) : "memory");
}

struct CycBlockDraw { uint32_t j0, j1; double v0, v1; };

__device__ __forceinline__ CycBlockDraw cyc_draw_block_v2(uint32_t& x, uint32_t kb, uint32_t b, double c1, double c0, bool advance) {
    CycBlockDraw r;
    uint32_t jit0, jit1;
#if CYC_KNOCK & 1   // no mixer
    const uint32_t w0 = x, w1 = x + kb;
#else
    const uint32_t w0 = lhd2_mix32(x), w1 = lhd2_mix32(x + kb);
#endif
#if CYC_KNOCK & 2   // no wide multiply
    r.j0 = (w0 >> 28) & (2u * b); jit0 = w0; r.j1 = (w1 >> 28) & (2u * b + 1u); jit1 = w1;
#else
    mulhilo32(w0, 2u * b + 1u, r.j0, jit0);
    mulhilo32(w1, 2u * b + 2u, r.j1, jit1);
#endif
    if (advance) x += 2u * kb;  // Weyl sequence A + i B
#if CYC_KNOCK & 4   // no DFMA in the value
    r.v0 = __hiloint2double((int)(0x3fe00000u | (2u * b)), (int)jit0);
    r.v1 = __hiloint2double((int)(0x3fe00000u | (2u * b + 1u)), (int)jit1);
#else
    r.v0 = lhd2_value(2u * b, jit0, c1, c0);
    r.v1 = lhd2_value(2u * b + 1u, jit1, c1, c0);
#endif
    return r;
}

template <int N, int D, int C>
__device__ __forceinline__ void cyc_generate_v2(unsigned int cand, uint32_t s_lo, uint32_t s_hi, double c1, double c0, int g,
                                                unsigned int group_mask) {
    static_assert(N % 2 == 0 && (D == 2 || D == 3), "row blocks of two; two lanes own two or three columns");
    constexpr int NB = N / 2, H = (NB + 1) / 2;
    constexpr unsigned RB = Layout<N, D, C>::kRowBytes, DB = Layout<N, D, C>::kDimBytes;
    uint32_t xp, pb, xs = 0u, sb = 0u;
    lhd2_column_key((uint32_t)g, s_lo, s_hi, xp, pb);
    if (D == 3) {
        lhd2_column_key(2u, s_lo, s_hi, xs, sb);
        if (g) xs += (uint32_t)(2 * H) * sb;  // lane 1 takes the third column over at block H
    }
    const unsigned int colp = cand + (unsigned int)g * DB, cols = cand + 2u * DB;
    auto phase = [&](int t_begin, int t_end, bool sec) {
#pragma unroll 2
        for (int t = t_begin; t < t_end; ++t) {
            const CycBlockDraw P = cyc_draw_block_v2(xp, pb, (uint32_t)t, c1, c0, true);
            CycBlockDraw S{};
            if (D == 3) S = cyc_draw_block_v2(xs, sb, (uint32_t)t, c1, c0, sec);
            const unsigned int a0 = (unsigned int)(2 * t) * RB;
            // both elements of both blocks: loads first
#if CYC_KNOCK & 8   // no shared-memory traffic: one store per block so that the draws stay live
            cyc_sts(colp + a0, P.v0 + P.v1 + (double)(P.j0 + P.j1), true);
            if (D == 3) cyc_sts(cols + a0, S.v0 + S.v1 + (double)(S.j0 + S.j1), sec);
            continue;
#endif
            double p0 = cyc_lds(colp + P.j0 * RB, true, 0.0), p1 = cyc_lds(colp + P.j1 * RB, true, 0.0);
            double q0 = 0.0, q1 = 0.0;
            if (D == 3) { q0 = cyc_lds(cols + S.j0 * RB, sec, 0.0); q1 = cyc_lds(cols + S.j1 * RB, sec, 0.0); }
            // what element 2t+1 reads AFTER the stores of element 2t
            if (P.j1 == P.j0) p1 = P.v0; else if (P.j1 == 2u * t) p1 = p0;
            if (D == 3) { if (S.j1 == S.j0) q1 = S.v0; else if (S.j1 == 2u * t) q1 = q0; }
            cyc_sts(colp + a0, p0, true);
            cyc_sts(colp + P.j0 * RB, P.v0, true);
            cyc_sts(colp + a0 + RB, p1, true);
            cyc_sts(colp + P.j1 * RB, P.v1, true);
            if (D == 3) {
                cyc_sts(cols + a0, q0, sec);
                cyc_sts(cols + S.j0 * RB, S.v0, sec);
                cyc_sts(cols + a0 + RB, q1, sec);
                cyc_sts(cols + S.j1 * RB, S.v1, sec);
            }
        }
    };
    if (D == 3) {
        phase(0, H, g == 0);
        __syncwarp(group_mask);  // the third column's first half is in place before lane 1 continues it
        phase(H, NB, g == 1);
    } else {
        phase(0, NB, false);
    }
}

// One tile of the cyclic schedule: R rows against the lane's KP shifts.  The R + KP - 1 partner rows
// are read once for all shifts (16 row loads per 60 pairs at R = 5, KP = 12).  Tiles whose partner
// rows stay below N for both lanes (WRAP = false) address them as base + immediate; only the
// upper tiles pay one add and one unsigned min per row for the wrap-around.
template <int N, int D, int C, bool MAXPRO, bool SAFE, bool WRAP, int R, int KP>
__device__ __forceinline__ void cyc_tile_window(const unsigned char* cand, unsigned i0_off, unsigned first_off,
                                                const double (&xi)[R][D], RecipChain (&ch)[R], double (&mins)[R], double& acc) {
    constexpr unsigned RB = Layout<N, D, C>::kRowBytes;
    constexpr unsigned NB = N * RB;
    constexpr int W = R + KP - 1;
    double xj[W][D];
#pragma unroll
    for (int u = 0; u < W; ++u) {
        unsigned off = i0_off + first_off + u * RB;
        if (WRAP) off = min(off, off - NB);
        load_row_at<N, D, C>(xj[u], cand, off);
    }
#pragma unroll
    for (int c = 0; c < KP; ++c) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (MAXPRO) { if (SAFE) acc += 1.0 / cyc_maxpro_term<D>(xi[r], xj[r + c]); else ch[r].push(cyc_maxpro_term<D>(xi[r], xj[r + c])); }
            else mins[r] = min_sqdist(mins[r], cyc_sqdist_term<D>(xi[r], xj[r + c]));
        }
    }
}

// SAFE: designs a caller supplied may lie outside the unit cube, where the reciprocal chains
// are not range-safe; then every term takes an IEEE divide (src/maxpro_utils.rs:54 literally).
//
// FP64 instructions per pair: 3 DADD + 2 DMUL + 1 DFMA for the term, 1 DFMA + 1 DMUL for the
// chain push = 8, plus 4 per chain flush.  The chains start at den = 2^1000 (common.cuh), so one
// flush serves three tiles (36-39 terms) instead of one: 20 flushes per lane and candidate
// instead of 50.  The n/2 diameters are split between the two lanes by ADDRESS (lane 0 rows
// 0..12, lane 1 rows 13..24), not by predicate: 13 pair slots per lane instead of 25.
template <int N, int D, int C, bool MAXPRO, bool SAFE = false>
__device__ __forceinline__ double cyc_score(const unsigned char* cand, int g) {
    constexpr int R = 5, G = 2;
    constexpr int kHalf = (N - 1) / 2;
    constexpr int kPerLane = kHalf / G;
    constexpr int kTiles = N / R;
    constexpr int kNoWrap = (N - kHalf - R) / R + 1;  // tiles with t*R + kHalf + R-1 <= N-1
    constexpr int kFlushEvery = 3;                    // tiles between flushes
    constexpr int kDiam = (N % 2 == 0) ? (N / 2 + 1) / 2 : 0;  // diameter pairs per lane
    static_assert(N % R == 0 && kHalf % G == 0, "whole tiles, equal shift ranges");
    static_assert(kFlushEvery * kPerLane + (kDiam + R - 1) / R <= kMaxChainTerms, "chains stay in range between flushes");
    constexpr unsigned RB = Layout<N, D, C>::kRowBytes;

    RecipChain ch[R];
    double mins[R];
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) { ch[r].reset(); mins[r] = CUDART_INF; }
    const unsigned first_off = (1u + (unsigned)g * kPerLane) * RB;  // this lane's first shift
    auto flush_all = [&]() {
        if (MAXPRO && !SAFE) {
#pragma unroll
            for (int r = 0; r < R; ++r) acc = ch[r].flush_into(acc);
        }
    };

#pragma unroll 1
    for (int t = 0; t < kNoWrap; ++t) {
        const unsigned i0_off = (unsigned)t * R * RB;
        double xi[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r) load_row_at<N, D, C>(xi[r], cand, i0_off + r * RB);
        cyc_tile_window<N, D, C, MAXPRO, SAFE, false, R, kPerLane>(cand, i0_off, first_off, xi, ch, mins, acc);
        if (t % kFlushEvery == kFlushEvery - 1) flush_all();
    }
    // diameters {i, i + N/2} (even N), i < N/2: lane 0 takes i = 0 .. kDiam-1, lane 1 the rest.  The
    // rows are addressed per lane, so both lanes run the same instructions (no divergence); when
    // N/2 is odd lane 1's last slot is a repeat of row 0 that is not pushed.
    if (N % 2 == 0) {
        constexpr int H = N / 2;
#pragma unroll
        for (int j = 0; j < kDiam; ++j) {
            const bool live = (H % 2 == 0) || j < kDiam - 1 || g == 0;
            const unsigned off = live ? ((unsigned)g * kDiam + j) * RB : 0u;
            double xa[D], xb[D];
            load_row_at<N, D, C>(xa, cand, off);
            load_row_at<N, D, C>(xb, cand, off + H * RB);
            if (MAXPRO) {
                const double pterm = cyc_maxpro_term<D>(xa, xb);
                if (live) { if (SAFE) acc += 1.0 / pterm; else ch[j % R].push(pterm); }
            } else {
                const double dterm = cyc_sqdist_term<D>(xa, xb);
                if (live) mins[j % R] = min_sqdist(mins[j % R], dterm);
            }
        }
    }
    flush_all();
#pragma unroll 1
    for (int t = kNoWrap; t < kTiles; ++t) {
        const unsigned i0_off = (unsigned)t * R * RB;
        double xi[R][D];
#pragma unroll
        for (int r = 0; r < R; ++r) load_row_at<N, D, C>(xi[r], cand, i0_off + r * RB);
        cyc_tile_window<N, D, C, MAXPRO, SAFE, true, R, kPerLane>(cand, i0_off, first_off, xi, ch, mins, acc);
        if ((t - kNoWrap) % kFlushEvery == kFlushEvery - 1) flush_all();
    }
    if ((kTiles - kNoWrap) % kFlushEvery != 0) flush_all();
    if (MAXPRO) return acc;
    double m = mins[0];
#pragma unroll
    for (int r = 1; r < R; ++r) m = min_sqdist(m, mins[r]);
    return m;
}

// Supplied designs: HBM row-major [i][k] -> shared [k][i][c].  Lane 0 takes the first kRows0 rows,
// lane 1 the rest, as 16-byte vectors (LDG.128: a 32-byte sector is fetched for every 16 bytes
// used, the best a lane-per-design access can do without staging).  Returns false if a value
// lies outside [0, 1] (sign bit set or high word above that of 1.0): the caller then scores
// the design with IEEE divides.
template <int N, int D, int C>
__device__ __forceinline__ bool cyc_load_design(unsigned char* cand, const double* src, int g) {
    constexpr int kElems = N * D;
    static_assert(kElems % 2 == 0, "16-byte vectors");
    constexpr int kRows0 = ((N / 2 + 1) / 2) * 2;  // lane 0's rows: even, so both lanes start on a 16-byte boundary
    constexpr int kVec0 = kRows0 * D / 2, kVec1 = (kElems - kRows0 * D) / 2;
    static_assert((kRows0 * D) % 2 == 0 && kVec1 >= 0 && kVec1 <= kVec0, "split");
    constexpr unsigned RB = Layout<N, D, C>::kRowBytes, DB = Layout<N, D, C>::kDimBytes;
    const double2* v = reinterpret_cast<const double2*>(src) + g * kVec0;
    unsigned char* dst = cand + (unsigned)g * kRows0 * RB;
    unsigned int worst = 0u;
#pragma unroll
    for (int t = 0; t < kVec0; ++t) {
        if (t < kVec1 || g == 0) {
            const double2 x = __ldg(v + t);
            const int e0 = 2 * t, e1 = 2 * t + 1;
            *reinterpret_cast<double*>(dst + (e0 / D) * RB + (e0 % D) * DB) = x.x;
            *reinterpret_cast<double*>(dst + (e1 / D) * RB + (e1 % D) * DB) = x.y;
            worst = max(worst, max((unsigned int)__double2hiint(x.x), (unsigned int)__double2hiint(x.y)));
            // exactly-1.0 passes (hi == 0x3FF00000, lo == 0); anything in (1, 1+2^-20) also passes and is harmless
        }
    }
    return worst <= 0x3FF00000u;
}

// Named-barrier token between two warps (64 threads): bar.arrive hands it on, bar.sync takes it.
__device__ __forceinline__ void cyc_token_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cyc_token_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// Turn-taking on the FP64 pipe.  Each warp alternates an integer-only, latency-bound phase
// (Philox + Fisher-Yates, ~15k cycles) and an FP64-bound phase (the pair loop, ~10k pipe
// cycles).  The warps of a sub-partition (warp id mod 4) share the pipe fairly, so left alone
// they tend to finish scoring together and then generate together.  With p.ring = 2 the first
// two warps of every sub-partition pass a token (strict alternation) and the third runs free;
// p.ring = 3 serialises all three.  MEASURED (profiles/README.md): neither helps -- 1.07 / 1.01
// vs 1.08 Gcand/s without.  The micro-probes of tools/micro (philox_lat.cu, issue2.cu) say why:
// Philox's IMAD.WIDE + LOP3 stream and a DFMA stream do not overlap on a sub-partition, whether
// they come from two warps or are interleaved in one (a Philox warp slows a DFMA warp beside it
// from 2.07 to 3.56 cycles per instruction), so the two phases cost their sum however they are
// arranged, and a lone scoring warp reaches only ~62% of the pipe.  Kept as an experiment switch
// (MAXPRO_CYCLIC_RING), off by default.
template <int N, int D, int C, bool MAXPRO, bool SUPPLIED = false, int GEN = 2>
__global__ void __launch_bounds__(2 * C, 1) search_cyclic_kernel(CyclicParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int G = 2, kPerWarp = 16;
    static_assert(N % 2 == 0 ? (N / 2) % 5 == 0 : true, "diameter tiles");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = wid * kPerWarp + (lane & 15), g = lane >> 4;
    const unsigned int group_mask = (1u << (lane & 15)) | (1u << ((lane & 15) + 16));
    unsigned char* cand = smem_raw + (size_t)c * 8;
    BestRec* warp_best = reinterpret_cast<BestRec*>(smem_raw + (size_t)N * D * C * sizeof(double));

    const unsigned long long gcid = (unsigned long long)blockIdx.x * C + c;
    const unsigned long long gstep = (unsigned long long)gridDim.x * C;
    const unsigned long long* list = p.list;
    unsigned long long list_count = p.list_count;
    if (list && p.list_count_dev) {
        list_count = *p.list_count_dev;
        if (list_count > p.list_cap) list = nullptr;  // stage 1 overflowed its list: the exact search over the whole range
    }
    const unsigned long long total = list ? list_count : p.hi - p.lo;
    const unsigned long long rounds = (total + gstep - 1) / gstep;

#ifdef CYC_PROF
    long long prof_lock = 0, prof_gen = 0, prof_score = 0, prof_start[3] = {0, 0, 0};
#endif
    LaneBest<MAXPRO> lb;  // fold of src/lib.rs:85-97 on the criterion (reduce.cuh)
    lb.init(p.inv_d);

    // token ring: barrier 1 + smsp*ring + q is the token INTO ring position q
    const int ring = p.ring, smsp = wid & 3, q = wid >> 2;
    const bool in_ring = q < ring;
    const int tok_in = 1 + smsp * ring + q, tok_out = 1 + smsp * ring + (q + 1 == ring ? 0 : q + 1);
    if (in_ring && q == ring - 1) cyc_token_pass(tok_out);  // position 0 scores first

    for (unsigned long long rnd = 0; rnd < rounds; ++rnd) {
        const unsigned long long off = gcid + rnd * gstep;
        bool active = off < total;
        unsigned long long idx = p.lo + off;
#ifdef CYC_PROF
        const long long pt0 = clock64();
#endif
#ifdef CYC_PROF
        const long long pt1 = clock64();
#endif
        if (list) {
            idx = active ? list[off] : kNoIndex;
            active = idx != kNoIndex;
        }
        const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
        const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
        bool in_cube = true;
        if (active && SUPPLIED) {
            in_cube = cyc_load_design<N, D, C>(cand, p.designs + idx * (unsigned long long)(N * D), g);
        } else if (active && (CYC_EXP == 7 || CYC_EXP == 8) && (q < 2 ? rnd > 0 : CYC_EXP == 7)) {
            // experiment: warps 0-1 of a sub-partition score only, warp 2 idles (7) or generates only (8)
        } else if (active && CYC_EXP == 8 && q == 2) {
#pragma unroll 1
            for (int rep = 0; rep < 3; ++rep)
                cyc_generate_v2<N, D, C>((unsigned int)__cvta_generic_to_shared(cand), s_lo + rep, s_hi, p.c1, p.c0, g, group_mask);
        } else if (active && CYC_EXP == 1 && rnd > 0) {
            // experiment: no stand-alone generation after the first round
        } else if (active) {
            if (GEN == 2 && N % 2 == 0 && !MAXPRO) {
                // side-by-side columns: measured +2.4 % for maximin, -2.4 % for MaxPro (issue-bound: more element slots)
                cyc_generate_v2<N, D, C>((unsigned int)__cvta_generic_to_shared(cand), s_lo, s_hi, p.c1, p.c0, g, group_mask);
            } else if (D == 3) {
                cyc_generate_three<N, C, GEN>(cand, s_lo, s_hi, p.c1, p.c0, g, group_mask);
            } else {
#pragma unroll 1
                for (int k = g; k < D; k += G)
                    cyc_generate_blocks<N, D, C, GEN>(cand + (size_t)k * Layout<N, D, C>::kDimBytes, (uint32_t)k, 0, (N + 1) / 2, s_lo, s_hi, p.c1, p.c0);
            }
        }
        __syncwarp(group_mask);
#ifdef CYC_PROF
        __syncwarp();
        const long long pt2 = clock64();
#endif
        if (in_ring) cyc_token_wait(tok_in);
        double s = 0.0;
        if (SUPPLIED && MAXPRO) {
            in_cube = __shfl_xor_sync(group_mask, (int)in_cube, 16) && in_cube;  // both halves of the design
            if (active && in_cube) s = cyc_score<N, D, C, MAXPRO, false>(cand, g);
            else if (active) s = cyc_score<N, D, C, MAXPRO, true>(cand, g);
        } else if (active) {
#if CYC_EXP == 3
            s = *reinterpret_cast<const double*>(cand + (s_lo % 150u) * Layout<N, D, C>::kRowBytes);
#elif CYC_EXP == 7 || CYC_EXP == 8
            if (q < 2) s = cyc_score<N, D, C, MAXPRO>(cand, g);
            else s = *reinterpret_cast<const double*>(cand + (s_lo % 150u) * Layout<N, D, C>::kRowBytes);
#else
            s = cyc_score<N, D, C, MAXPRO>(cand, g);
#endif
        }
        if (in_ring) cyc_token_pass(tok_out);
#ifdef CYC_PROF
        __syncwarp();
        { const long long pt3 = clock64(); prof_lock += pt1 - pt0; prof_gen += pt2 - pt1; prof_score += pt3 - pt2;
          if (rnd == 100 || rnd == 101 || rnd == 102) prof_start[rnd - 100] = pt2; }
#endif
        const double other = __shfl_xor_sync(group_mask, s, 16);
        if (MAXPRO) s = (g == 0) ? s + other : other + s;  // same bits in both lanes
        else s = fmin(s, other);
        __syncwarp(group_mask);
        if (!MAXPRO) s = sqrt(s);
        if (SUPPLIED) {
            // every score leaves the chip
            if (MAXPRO) s = maxpro_psi(s, p.n_pairs, p.inv_d);  // maxpro_utils.rs:81-82
            if (active && g == 0) p.scores[idx] = s;
            if (active) lb.offer_criterion(s, idx);
        } else if (active) {
            lb.offer(s, idx, p.n_pairs, p.inv_d);
        }
    }
#ifdef CYC_PROF
    if (blockIdx.x == 0 && lane == 0)
        printf("warp %2d smsp %d q %d rounds %llu: lock %lld gen %lld score %lld cycles/round; score starts of rounds 100-102: %lld %lld %lld\n", wid, smsp, q, rounds,
               prof_lock / (long long)rounds, prof_gen / (long long)rounds, prof_score / (long long)rounds, prof_start[0] % 10000000, prof_start[1] % 10000000, prof_start[2] % 10000000);
#endif
    if (in_ring && q == 0) cyc_token_wait(tok_in);  // take the last token so no arrival is left pending
    block_publish_best<MAXPRO>(lb.best, lb.idx, warp_best, p.block_best, p.done_counter, p.result);
}

template <int N, int D, int C>
static cudaError_t launch_cyclic_nd(const SearchLaunch& L, const CyclicParams& p) {
    constexpr size_t smem = (size_t)N * D * C * sizeof(double) + 32 * sizeof(BestRec);
    static_assert(smem <= kSmemBudget, "candidates per CTA do not fit shared memory");
    cudaError_t e;
    auto go = [&](auto kernel) -> cudaError_t {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<L.grid, 2 * C, smem, L.stream>>>(p);
        return cudaGetLastError();
    };
    if (p.designs) return L.metric == 0 ? go(search_cyclic_kernel<N, D, C, true, true>) : go(search_cyclic_kernel<N, D, C, false, true>);
    if (p.gen == 1) return L.metric == 0 ? go(search_cyclic_kernel<N, D, C, true, false, 1>) : go(search_cyclic_kernel<N, D, C, false, false, 1>);
    return L.metric == 0 ? go(search_cyclic_kernel<N, D, C, true, false, 2>) : go(search_cyclic_kernel<N, D, C, false, false, 2>);
}

// Shapes with a specialised instantiation; *cands = candidates per CTA.
bool search_cyclic_supported(unsigned long long n, unsigned long long d, int* cands) {
    if (const char* e = cfg("MAXPRO_NO_CYCLIC")) if (atoi(e) != 0) return false;
    int c = 0;
    if (n == 50 && d == 3) c = 192;
    else if (n == 50 && d == 2) c = 256;
    if (cands) *cands = c;
    return c != 0;
}

cudaError_t launch_search_cyclic(const SearchLaunch& L) {
    CyclicParams p;
    p.ring = 0;  // measured: turn-taking does not pay (generation and scoring cost their sum either way)
    if (const char* e = cfg("MAXPRO_CYCLIC_RING")) { int v = atoi(e); if (v == 0 || v == 2 || v == 3) p.ring = v; }
    if (p.ring > (L.threads / 32) / 4) p.ring = 0;
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    p.list = L.list; p.list_count = L.list_count; p.list_count_dev = L.list_count_dev; p.list_cap = L.list_cap;
    p.designs = L.designs; p.scores = L.scores;
    p.gen = stream_version();
    p.c1 = ldexp(1.0 / (double)L.n, -32);
    p.c0 = lhd_c0(p.gen, p.c1);
    p.n_pairs = (double)L.n * ((double)L.n - 1.0) / 2.0;
    p.inv_d = 1.0 / (double)L.d;
    p.block_best = L.block_best; p.done_counter = L.done_counter; p.result = L.result;
    if (L.n == 50 && L.d == 3) return launch_cyclic_nd<50, 3, 192>(L, p);
    if (L.n == 50 && L.d == 2) return launch_cyclic_nd<50, 2, 256>(L, p);
    return cudaErrorInvalidValue;
}

}  // namespace mp

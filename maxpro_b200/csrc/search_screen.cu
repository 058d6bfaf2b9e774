// K1+K2 fused, TWO-STAGE search for small designs, both metrics (a kernel specialised for the headline shapes n = 50,
// d = 2 or 3, and a general one for any n from 8 to 256 with d up to 10 further down): an fp32 screening
// pass with a PROVEN error bound, then the exact fp64 kernel on the survivors.  The result is the result of the
// exact search -- same winner under the reference's fold (src/lib.rs:85-97, criterion after powf, later index on
// ties), same score -- at about twice its rate.
//
// north_star: "fp32-product / fp64-accumulate variant benchmarked separately"; SURVEY.md 8f.1: a two-stage search
// that "keeps 1e-9 parity on the reported winner".  It replaces the same reference code as the fp64 kernels
// (src/lib.rs:65-98, src/lhd.rs:97-132, src/maxpro_utils.rs:28-83) and draws the same design stream.
//
//   stage 1  (this file) every candidate of the range is generated and scored in fp32.  A candidate SURVIVES if
//            its fp32 score is not above a threshold derived from the best true upper bound seen so far by its
//            CTA; the threshold is wide enough that the true winner -- and every candidate whose criterion can
//            tie with it -- is certain to be below it (the bound is derived below).  Survivors are appended to a
//            list in HBM (8 bytes each; a few hundred per CTA for 2^28 candidates).
//   stage 2  the fp64 kernel of the shape (search_cyclic.cu, search_small.cu, search_warp.cu) in list mode regenerates
//            the survivors, scores them in fp64 and folds them exactly as the full fp64 search would.  If the list overflows, stage 2 scores the whole range instead.
//
// ---- fp32 arithmetic with exact differences ----------------------------------------------------------------
// A coordinate of the design stream is x = (i + (jit + 1/2) 2^-32) / n.  Stage 1 keeps y = i 2^18 + (jit >> 14), an
// INTEGER below 2^24 (n <= 64), as a float: exactly representable, and so is every difference y_a - y_b.  The
// only departure from the true design is the truncated jitter: |x_a - x_b - (y_a - y_b) / (n 2^18)| < eta =
// 2^-18 / n.  Column 0 carries an exact power-of-two factor so that the scaled terms
//     p_s = q_s^2 + eps_s,   q_s = G prod_k (x_ak - x_bk),   eps_s = G^2 eps        (G: exact scale, G^2 eps ~ 2^-20)
// lie in [2^-20, 2^20]: six of them can be chained ((num, den) <- (num p + den, den p)) inside the fp32 range, and
// S = G^2 sum 1/p_s.  Rounding happens in two products, one fma, the chain and the sums: below 2^-18 relative in
// total (`arith`).
//
// ---- the bound -----------------------------------------------------------------------------------------------
// With |delta'_k - delta_k| < eta and |delta_k| <= 1: |q' - q| <= dq = d eta (1 + eta)^(d-1).  For one term
// f = 1 / (q^2 + eps) the computed f' = 1 / (q'^2 + eps) satisfies
//     f'/f >= (q^2 + eps) / ((|q| + dq)^2 + eps) >= gmin                      for every q  (gmin ~ 0.75: host, numerically)
//     |f'/f - 1| <= rho(F) = (2 Q dq + dq^2) / (1/F - 2 Q dq),  Q = sqrt(1/F - eps)   whenever f <= F < 1/(2 eps)
// (the ratio is worst at the smallest admissible |q|, which is Q).  Every term of a design is at most its sum S, so
//   (i)  S <= S'/gmin =: F0, hence every term is <= F0 and S <= U := S' / (1 - rho(F0)) (1 + arith):
//        U is a TRUE upper bound of that candidate's S, and so of the minimum over the range;
//   (ii) a candidate with S <= U (1 + tau) has S' <= W := U (1 + tau) (1 + rho(U (1 + tau))) (1 + arith),
//        where tau = (4d + 4) 2^-52 is the band inside which two sums can share a criterion (reduce.cuh).
// The winner and everything that can tie with it have S <= U (1 + tau) for the U of ANY candidate, so keeping every
// candidate with S' <= W, for the smallest W known when it is scored, loses none of them.  W / S' is about 1.07
// at n = 50, d = 3: the survivors are the candidates within a few percent of the running best.
//
// ---- a second, per-candidate bound (MaxPro) ---------------------------------------------------------------------------
// The bound above charges every term with the worst case "one difference changed by eta, the others equal to 1".  In a
// Latin hypercube much more is known: column k, sorted, runs through the strata in order, so no two of its values are
// closer than the smallest gap between ADJACENT strata, and generation sees those gaps as it goes (stratum i is drawn
// right after stratum i-1).  With m the smallest adjacent gap of the candidate over all columns, in the integer units of y,
// every computed difference has |d'| >= m and the true one lies within one unit of it, so for every pair
//     q / q' = prod_k (1 + e_k / d'_k)   lies in   [(1 - 1/m)^d, (1 + 1/m)^d],      theta := (1 + 1/m)^d - 1,
// and f'/f = (q^2 + eps) / (q'^2 + eps) lies between 1 and q^2/q'^2, hence in [(1 - theta)^2, (1 + theta)^2]: every term,
// and so the sum, is off by at most r = 2 theta + theta^2 relative -- whatever d is and however much eps matters.
// m is a few thousand units at least for all but a fraction of a percent of the candidates (r ~ 1e-3 and less), so
//     S'/((1 + r)(1 + arith)) =: L  <=  S  <=  U := S' (1 + arith)/(1 - r)
// is far tighter than the first bound; a candidate survives only if its L is not above the smallest U its CTA has seen
// (times the tie band).  Both tests are necessary conditions for being the winner, so both are applied; the first one
// alone stopped paying at d = 5 (W/S' too wide), the second works up to the d = 8 the kernels are built for.
//
// ---- maximin (same two stages) ----------------------------------------------------------------------------------
// The criterion is the smallest pairwise distance, larger is better (src/maximin_utils.rs:54-67).  Stage 1 takes the
// minimum m' over the pairs of s' = sum_k (y_ak - y_bk)^2 in fp32, in units of u = n 2^18 (no scaling: s' < 2^50).
// Each true difference lies within one unit of the exact integer difference, so for every pair
//     |s u^2 - s'| <= E(s') = 2 sqrt(d) sqrt(s') + d + arith s'        (Cauchy-Schwarz; arith: the d roundings of the sum)
// and, E being increasing while s' - E(s') is increasing above a few units, the true minimum m satisfies
//     lower(m') = m' - E(m') <= m u^2 <= m' + E(m') = upper(m').
// A candidate survives if upper(m') is not below the largest lower(.) its CTA has seen: the winner and everything that
// ties with it (sqrt can map neighbouring minima to one criterion) pass, about one candidate in 10^5 at n = 50, d = 3.
#include "common.cuh"
#include "philox.cuh"
#include "reduce.cuh"
#include "kernels.h"

#include <cmath>

#ifndef SCR_EXP
#define SCR_EXP 0
#endif
#ifndef SCR_KNOCK
#define SCR_KNOCK 0  // experiment: parts of the generator knocked out (bit mask, see the kernel)
#endif
#ifndef SCR_TILE
#define SCR_TILE 4  // row blocks per scoring tile (measured: profiles/README.md, round 2)
#endif

namespace mp {

struct ScreenBound {
    double g2;      // S_true = g2 * S_scaled
    double eps;     // the reference's eps as stage 1 sees it: eps_s / g2 (eps_s is a rounded float)
    double dq;      // bound of |q' - q|
    double gmin;    // lower bound of f'/f over all q
    double arith;   // relative bound of all fp32 rounding in one score
    double tau;     // criterion tie band
    double mm_c1, mm_c0, mm_arith;  // maximin: E(s') = mm_c1 sqrt(s') + mm_c0 + mm_arith s'
};

struct ScreenParams {
    int gen;                         // design stream version (philox.cuh)
    int n, d;                        // points and dimensions of a design (the general kernel reads n; both read d for the bound)
    int group;                       // general kernel: lanes per candidate (1, or 2 = one role per half-warp)
    int jbits;                       // jitter bits kept by stage 1: y = i 2^jbits + (jit >> (32 - jbits)) < 2^24 (18 up to n = 64)
    unsigned long long seed, lo, hi;
    float col0_scale;                // exact power of two carried by column 0 (d <= 5) ...
    float col_scale[10];             // ... or spread over all columns (d >= 6: one column cannot hold 2^-130 and more)
    float eps_s;                     // eps in scaled units
    ScreenBound bound;
    unsigned long long* list;        // survivors (candidate indices)
    unsigned int* count;             // appended so far (may exceed cap: overflow)
    unsigned int cap;
    double* debug_scores;            // tests: scaled fp32 score of candidate lo + t at [t] (null otherwise)
};

// rho(F) of the header comment; returns a negative value where the bound does not apply
__host__ __device__ inline double screen_rho(double F, const ScreenBound& b) {
    const double inv = 1.0 / F;
    if (!(inv > 2.0 * b.eps)) return -1.0;
    const double Q = sqrt(inv - b.eps);
    const double den = inv - 2.0 * Q * b.dq;
    if (!(den > 0.0)) return -1.0;
    return (2.0 * Q * b.dq + b.dq * b.dq) / den;
}

// W of the header comment for a candidate whose scaled fp32 score is s; +inf where no bound applies
__host__ __device__ inline double screen_threshold(double s, const ScreenBound& b) {
    const double inf = 1.0 / 0.0;
    const double s_true = s * b.g2;
    const double r0 = screen_rho(s_true / b.gmin, b);
    if (r0 < 0.0 || r0 >= 0.5) return inf;
    const double U = s_true / (1.0 - r0) * (1.0 + b.arith) * (1.0 + b.tau);
    const double r1 = screen_rho(U, b);
    if (r1 < 0.0) return inf;
    return U * (1.0 + r1) * (1.0 + b.arith) / b.g2;
}

// maximin: bounds of the true squared minimum (units u^2) around the fp32 minimum m
__host__ __device__ inline double screen_mm_err(double m, const ScreenBound& b) { return b.mm_c1 * sqrt(m) + b.mm_c0 + b.mm_arith * m; }
__host__ __device__ inline double screen_mm_lower(double m, const ScreenBound& b) { const double l = m - screen_mm_err(m, b); return l > 0.0 ? l : 0.0; }
__host__ __device__ inline double screen_mm_upper(double m, const ScreenBound& b) { return m + screen_mm_err(m, b); }

namespace {

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 do two lanes' worth per issue slot) -----------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ f2 lds2(unsigned int addr) { f2 r; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(addr)); return r; }

// sum of reciprocals of up to six packed terms without a divide per term (see header): first term starts the chain
struct Chain2 {
    f2 num, den;
    __device__ __forceinline__ void start(f2 p) { num = pk(1.0f, 1.0f); den = p; }  // num is folded into the first push by the compiler
    __device__ __forceinline__ void push(f2 p) { num = fma2(num, p, den); den = mul2(den, p); }
    __device__ __forceinline__ f2 value() const {
        float d0, d1;
        unpk(den, d0, d1);
        return mul2(num, pk(rcp_approx(d0), rcp_approx(d1)));
    }
};

// Score of T consecutive row blocks m .. m+T-1 (2T rows) against their partners, in scaled units: the sum of
// 1 / p_s over those pairs (see the header).  Block b is paired with the SH = (NB-1)/2 blocks that follow it
// (cyclically), which covers every pair of blocks once.  For a tile that is
//   * the blocks after it INSIDE the tile (already in registers),
//   * b "late" blocks m+SH+1 .. m+SH+b,
//   * the SH-T+1 blocks m+T .. m+SH, which are partners of every block of the tile: one read of such a block
//     from shared memory serves 4T pairs (T = 1: 4 pairs per read -- the first version of this kernel).
// Every row meets SH packed terms; with SH = 12 that is two reciprocal chains of six per row (the chain switch
// falls on the same common partner for all rows because every block has T-1 terms before its first common one).
// Block layout: block m of dimension k of candidate c (= lane) at ((k * NB + m) * C + c) * 8, the even row in the
// low half: one LDS.64, conflict-free across the warp.
template <int NB, int D, int C, int T>
__device__ __forceinline__ float screen_tile(unsigned int base, unsigned int m, f2 eps2) {
    constexpr int SH = (NB - 1) / 2, kChainTerms = 6;
    constexpr unsigned BS = C * 8u, DS = NB * BS;
    static_assert(SH + 1 > kChainTerms && SH <= 2 * kChainTerms && T - 1 < kChainTerms && T <= SH, "two chains of at most six terms per row");
    auto block_off = [&](unsigned int mm) { const unsigned int o = mm * BS; return min(o, o - DS); };  // block index modulo NB (mm < 2 NB)
    auto load_block = [&](f2 (&b)[D], unsigned int off) {
#pragma unroll
        for (int k = 0; k < D; ++k) b[k] = lds2(base + k * DS + off);
    };
    auto term2 = [&](const f2 (&a)[D], const f2 (&b)[D]) {
        f2 q = sub2(a[0], b[0]);
#pragma unroll
        for (int k = 1; k < D; ++k) q = mul2(q, sub2(a[k], b[k]));
        return fma2(q, q, eps2);
    };
    f2 A[T][D], X[T][2][D];
    float inner = 0.0f;
#pragma unroll
    for (int b = 0; b < T; ++b) {
        load_block(A[b], (m + (unsigned int)b) * BS);  // m + T - 1 < NB: no wrap
        float q_in = 1.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float a0, a1;
            unpk(A[b][k], a0, a1);
            X[b][0][k] = pk(a0, a0); X[b][1][k] = pk(a1, a1);
            q_in = k == 0 ? a0 - a1 : q_in * (a0 - a1);  // the pair inside the block
        }
        float e0, e1;
        unpk(eps2, e0, e1);
        inner += rcp_approx(fmaf(q_in, q_in, e0));
    }
    Chain2 ch[T][2];
    int cnt[T];  // terms a row of block b has met (compile-time after unrolling)
#pragma unroll
    for (int b = 0; b < T; ++b) cnt[b] = 0;
    f2 vsum = pk(0.0f, 0.0f);
    auto feed = [&](int b, const f2 (&P)[D]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const f2 pr = term2(X[b][r], P);
            if (cnt[b] == kChainTerms) vsum = add2(vsum, ch[b][r].value());
            if (cnt[b] == 0 || cnt[b] == kChainTerms) ch[b][r].start(pr); else ch[b][r].push(pr);
        }
        ++cnt[b];
    };
    // partners inside the tile
#pragma unroll
    for (int b = 0; b < T; ++b)
#pragma unroll
        for (int b2 = b + 1; b2 < T; ++b2) feed(b, A[b2]);
    // late partners: block m+SH+e is a partner of the blocks b >= e
#pragma unroll
    for (int e = 1; e < T; ++e) {
        f2 E[D];
        load_block(E, block_off(m + (unsigned int)(SH + e)));
#pragma unroll
        for (int b = e; b < T; ++b) feed(b, E);
    }
    // common partners
#pragma unroll
    for (int c = 0; c <= SH - T; ++c) {
        f2 B[D];
        load_block(B, block_off(m + (unsigned int)(T + c)));
#pragma unroll
        for (int b = 0; b < T; ++b) feed(b, B);
    }
#pragma unroll
    for (int b = 0; b < T; ++b) vsum = add2(vsum, add2(ch[b][0].value(), ch[b][1].value()));
    float v0, v1;
    unpk(vsum, v0, v1);
    return v0 + v1 + inner;
}

// The same tile walk for maximin: the smallest s' = sum_k (y_ak - y_bk)^2 over the tile's pairs (units u^2).
// Packed differences and sums as above; the running minimum takes both halves of a packed term in one FMNMX3.
template <int NB, int D, int C, int T>
__device__ __forceinline__ float screen_tile_min(unsigned int base, unsigned int m) {
    constexpr int SH = (NB - 1) / 2;
    constexpr unsigned BS = C * 8u, DS = NB * BS;
    static_assert(T <= SH, "tile within the shift range");
    auto block_off = [&](unsigned int mm) { const unsigned int o = mm * BS; return min(o, o - DS); };
    auto load_block = [&](f2 (&b)[D], unsigned int off) {
#pragma unroll
        for (int k = 0; k < D; ++k) b[k] = lds2(base + k * DS + off);
    };
    f2 A[T][D], X[T][2][D];
    float mn[T][2];
#pragma unroll
    for (int b = 0; b < T; ++b) {
        load_block(A[b], (m + (unsigned int)b) * BS);
        float s_in = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float a0, a1;
            unpk(A[b][k], a0, a1);
            X[b][0][k] = pk(a0, a0); X[b][1][k] = pk(a1, a1);
            s_in = k == 0 ? (a0 - a1) * (a0 - a1) : fmaf(a0 - a1, a0 - a1, s_in);  // the pair inside the block
        }
        mn[b][0] = s_in; mn[b][1] = s_in;
    }
    auto feed = [&](int b, const f2 (&P)[D]) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const f2 d0 = sub2(X[b][r][0], P[0]);
            f2 q = mul2(d0, d0);
#pragma unroll
            for (int k = 1; k < D; ++k) { const f2 dk = sub2(X[b][r][k], P[k]); q = fma2(dk, dk, q); }
            float lo, hi;
            unpk(q, lo, hi);
            mn[b][r] = min3(mn[b][r], lo, hi);
        }
    };
#pragma unroll
    for (int b = 0; b < T; ++b)
#pragma unroll
        for (int b2 = b + 1; b2 < T; ++b2) feed(b, A[b2]);
#pragma unroll
    for (int e = 1; e < T; ++e) {
        f2 E[D];
        load_block(E, block_off(m + (unsigned int)(SH + e)));
#pragma unroll
        for (int b = e; b < T; ++b) feed(b, E);
    }
#pragma unroll
    for (int c = 0; c <= SH - T; ++c) {
        f2 B[D];
        load_block(B, block_off(m + (unsigned int)(T + c)));
#pragma unroll
        for (int b = 0; b < T; ++b) feed(b, B);
    }
    float r = fminf(mn[0][0], mn[0][1]);
#pragma unroll
    for (int b = 1; b < T; ++b) r = min3(r, mn[b][0], mn[b][1]);
    return r;
}

// Generation of one candidate by one lane: the design stream's draws and shuffle (philox.cuh), values as exact
// integer-valued floats (see the header).  NB, BS and DS are compile-time constants in the specialised kernel and run-time
// values in the general one.  The D columns are independent Fisher-Yates chains (each step reads what an earlier step of the
// SAME column wrote), so they are walked side by side: D loads in flight per step instead of one.
template <int D, int GEN, bool ALLCOLS = false, bool TRACK = true>
__device__ __forceinline__ unsigned int screen_generate(unsigned int base, int NB, unsigned int BS, unsigned int DS, uint32_t s_lo, uint32_t s_hi,
                                                float col0_scale, int jbits, int k0 = 0, int kstep = 1, int dtot = D, bool single_last = false,
                                                const float* col_scale = nullptr) {
    // single_last: the design has an odd number of points -- the last block holds one element (its second half is never used)
    // this lane's columns: k0, k0 + kstep, ... below dtot (all D of them when one lane builds the whole design; the
    // general kernel with two lanes per candidate gives each lane every other column and masks the one that may be missing)
    uint32_t kb[D], x[D];
    int kc[D];
    bool on[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { kc[k] = k0 + k * kstep; on[k] = kc[k] < dtot; }
    if (GEN == 2) {
#pragma unroll
        for (int k = 0; k < D; ++k) lhd2_column_key((uint32_t)kc[k], s_lo, s_hi, x[k], kb[k]);
    }
    // smallest gap between adjacent strata over this lane's columns, in units of y (see the header: the second bound)
    unsigned int gap_min = 0xFFFFFFFFu, prev[D];
    const unsigned int one = 1u << jbits;
#pragma unroll 2
    for (int b = 0; b < NB; ++b) {
        uint32_t jx[D][2], jit[D][2];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (GEN == 1) {
                const Philox4 w = philox4x32_10((uint32_t)b, (uint32_t)kc[k], s_lo, s_hi, kKeyLhd0, kKeyLhd1);
                jit[k][0] = w.x; jx[k][0] = __umulhi(w.y, 2u * b + 1u);
                jit[k][1] = w.z; jx[k][1] = __umulhi(w.w, 2u * b + 2u);
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#if SCR_KNOCK & 1   // experiment: no mixer
                    const uint32_t w = x[k];
#else
                    const uint32_t w = lhd2_mix32(x[k]);
#endif
#if SCR_KNOCK & 2   // experiment: no wide multiply
                    jx[k][e] = (w >> 28) & (2u * b + e); jit[k][e] = w;
#else
                    mulhilo32(w, 2u * b + e + 1u, jx[k][e], jit[k][e]);
#endif
                    x[k] += kb[k];  // Weyl sequence A + i B
                }
            }
        }
        // both elements of the block in all D columns: loads first, then the stores in shuffle order.  What
        // element 2b+1 would have read after the stores of element 2b is patched in from registers (it can
        // only alias a[2b] or a[j0]).
        float t0[D], t1[D], v0[D], v1[D];
        unsigned int aj0[D], aj1[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const unsigned int f0 = jit[k][0] >> (32 - jbits), f1 = jit[k][1] >> (32 - jbits);
            v0[k] = (float)(((2u * b) << jbits) | f0);
            v1[k] = (float)(((2u * b + 1u) << jbits) | f1);
            if (TRACK && on[k]) {
                if (b > 0) gap_min = min(gap_min, one + f0 - prev[k]);                                  // strata 2b-1, 2b
                if (!(single_last && b == NB - 1)) gap_min = min(gap_min, one + f1 - f0);                // strata 2b, 2b+1
                prev[k] = f1;
            }
            if (ALLCOLS) { const float sc = col_scale[on[k] ? kc[k] : 0]; v0[k] *= sc; v1[k] *= sc; }
            else if (kc[k] == 0) { v0[k] *= col0_scale; v1[k] *= col0_scale; }
            aj0[k] = base + kc[k] * DS + (jx[k][0] >> 1) * BS + (jx[k][0] & 1u) * 4u;
            aj1[k] = base + kc[k] * DS + (jx[k][1] >> 1) * BS + (jx[k][1] & 1u) * 4u;
#if SCR_KNOCK & 8   // experiment: no shared-memory walk (one store per column block keeps the draws live)
            t0[k] = v0[k] + (float)aj0[k]; t1[k] = v1[k] + (float)aj1[k];
#else
            t0[k] = 0.0f; t1[k] = 0.0f;
            if (on[k]) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t0[k]) : "r"(aj0[k]) : "memory");
                if (!(single_last && b == NB - 1)) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t1[k]) : "r"(aj1[k]) : "memory");
            }
#endif
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (jx[k][1] == jx[k][0]) t1[k] = v0[k]; else if (jx[k][1] == 2u * b) t1[k] = t0[k];
            const unsigned int a0 = base + kc[k] * DS + (unsigned int)b * BS;
            if (!on[k]) continue;
#if SCR_KNOCK & 8
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0), "f"(t0[k] + t1[k]) : "memory");
#else
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0), "f"(t0[k]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(aj0[k]), "f"(v0[k]) : "memory");
            if (single_last && b == NB - 1) continue;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u), "f"(t1[k]) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(aj1[k]), "f"(v1[k]) : "memory");
#endif
        }
    }
    return TRACK ? gap_min : 0u;  // 0: no gap information, the second bound is not applied
}

// Survivor test of one candidate against its CTA's running bound (see the header), and the update of that bound.
// cta_thr holds bits of a non-negative double: MaxPro -- the smallest threshold W a candidate of this CTA implies;
// maximin -- the largest lower bound of a candidate's squared minimum.  Called by every thread of the CTA (rnd == 0 has
// a barrier: everybody's bound first, then the test against the best of them).
template <bool MAXPRO>
__device__ __forceinline__ void screen_survivors(const ScreenParams& p, unsigned long long* cta_thr, bool first_round, bool active,
                                                 double s, unsigned long long idx, unsigned int gap_min) {
    auto append = [&]() {
        const unsigned int pos = atomicAdd(p.count, 1u);
        if (pos < p.cap) p.list[pos] = idx;
    };
    const bool listed = SCR_EXP == 0 || (idx & 0xFFFFF) == 0;
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    if (MAXPRO) {
        // first bound: threshold W implied by this candidate's sum; second bound: [L, U] from its smallest adjacent gap
        double r = inf;  // gap_min == 0: the kernel does not track the gaps for this shape (the first bound is tight enough there)
        if (gap_min != 0u) {
            const double t = 1.0 + 1.0 / (double)gap_min;
            double pw = t;
            for (int k = 1; k < p.d; ++k) pw *= t;
            const double theta = pw - 1.0;
            r = theta * (2.0 + theta);
        }
        const double low = r < inf ? s / ((1.0 + r) * (1.0 + p.bound.arith)) : 0.0;
        const double up = r < 1.0 ? s * (1.0 + p.bound.arith) / (1.0 - r) * (1.0 + p.bound.tau) : inf;  // tie band folded in
        if (first_round) {
            if (active) {
                atomicMin(cta_thr, (unsigned long long)__double_as_longlong(screen_threshold(s, p.bound)));
                atomicMin(cta_thr + 1, (unsigned long long)__double_as_longlong(up));
            }
            __syncthreads();
            const double thr = __longlong_as_double((long long)cta_thr[0]), thr_u = __longlong_as_double((long long)cta_thr[1]);
            if (active && s <= thr && low <= thr_u && listed) append();
        } else {
            const double thr = __longlong_as_double((long long)*(volatile unsigned long long*)cta_thr);
            const double thr_u = __longlong_as_double((long long)*(volatile unsigned long long*)(cta_thr + 1));
            if (active && s <= thr && low <= thr_u && listed) {
                append();
                const double w = screen_threshold(s, p.bound);
                if (w < thr) atomicMin(cta_thr, (unsigned long long)__double_as_longlong(w));
                if (up < thr_u) atomicMin(cta_thr + 1, (unsigned long long)__double_as_longlong(up));
            }
        }
    } else {
        const double up = screen_mm_upper(s, p.bound), low = screen_mm_lower(s, p.bound);
        if (first_round) {
            if (active) atomicMax(cta_thr, (unsigned long long)__double_as_longlong(low));
            __syncthreads();
            const double thr = __longlong_as_double((long long)*cta_thr);
            if (active && up >= thr && listed) append();
        } else {
            const double thr = __longlong_as_double((long long)*(volatile unsigned long long*)cta_thr);
            if (active && up >= thr && listed) {
                append();
                if (low > thr) atomicMax(cta_thr, (unsigned long long)__double_as_longlong(low));
            }
        }
    }
}

// Shared-memory layout: rows in blocks of two, block m of dimension k of candidate c (= lane) at
// ((k * NB + m) * C + c) * 8, the even row in the low half.  A block is one LDS.64, conflict-free across the warp.
template <int N, int D, int C, int GEN, bool MAXPRO>
__global__ void __launch_bounds__(C, 1) search_screen_kernel(ScreenParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // MaxPro: [0] smallest survivor threshold W implied by a candidate of this CTA, [1] smallest upper bound U of a candidate's
    // sum; maximin: [0] largest lower bound of a candidate's squared minimum (bits of non-negative doubles: the integer
    // order is the numeric order)
    __shared__ unsigned long long cta_thr[2];
    static_assert(N % 2 == 0 && (N / 2) % 2 == 1 && N <= 64, "row blocks: an odd number of them, strata below 2^6");
    constexpr int NB = N / 2;  // row blocks of two
    constexpr unsigned BS = C * 8u, DS = NB * BS;
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(smem_raw) + threadIdx.x * 8u;

    const unsigned long long gtid = (unsigned long long)blockIdx.x * C + threadIdx.x;
    const unsigned long long gstep = (unsigned long long)gridDim.x * C;
    const unsigned long long total = p.hi - p.lo;
    const unsigned long long rounds = (total + gstep - 1) / gstep;
    if (threadIdx.x == 0) { cta_thr[0] = MAXPRO ? 0x7FF0000000000000ull : 0ull; cta_thr[1] = 0x7FF0000000000000ull; }
    __syncthreads();
    const f2 eps2 = pk(p.eps_s, p.eps_s);

    for (unsigned long long rnd = 0; rnd < rounds; ++rnd) {
        const unsigned long long off = gtid + rnd * gstep;
        const bool active = off < total;
        const unsigned long long idx = p.lo + off;
        double s = 0.0;
        unsigned int gap_min = 0u;
        if (active) {
            const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
            const uint32_t s_lo = (uint32_t)stream, s_hi = (uint32_t)(stream >> 32);
            if (!(SCR_EXP == 1 && rnd > 0))  // experiment 1: scoring only after the first round
                gap_min = screen_generate<D, GEN, false, false>(base, NB, BS, DS, s_lo, s_hi, p.col0_scale, 18);  // first bound only: tight at these shapes
            // ---- scoring: row blocks in tiles of kTile, the odd blocks out one by one (screen_tile above)
            constexpr int kTile = SCR_TILE;
            constexpr int kTiles = NB / kTile;
            if (MAXPRO) {
                double acc = 0.0;
#pragma unroll 1
                for (int t = 0; t < (SCR_EXP == 3 ? 1 : kTiles); ++t)
                    acc += (double)screen_tile<NB, D, C, kTile>(base, (unsigned int)(t * kTile), eps2);
                if (SCR_EXP != 3) {
#pragma unroll
                    for (int m = kTiles * kTile; m < NB; ++m) acc += (double)screen_tile<NB, D, C, 1>(base, (unsigned int)m, eps2);
                }
                s = acc;
            } else {
                float mn = CUDART_INF_F;
#pragma unroll 1
                for (int t = 0; t < kTiles; ++t) mn = fminf(mn, screen_tile_min<NB, D, C, kTile>(base, (unsigned int)(t * kTile)));
#pragma unroll
                for (int m = kTiles * kTile; m < NB; ++m) mn = fminf(mn, screen_tile_min<NB, D, C, 1>(base, (unsigned int)m));
                s = (double)mn;
            }
            if (p.debug_scores) p.debug_scores[off] = s;
        }
        screen_survivors<MAXPRO>(p, cta_thr, rnd == 0, active, s, idx, gap_min);
    }
}

// ---- the general kernel: any n with 8 <= n <= 256 (18 jitter bits up to n = 64, one fewer per doubling), D as a template parameter --------------------------------------
// Same two stages, same bound, same layout (block m of dimension k of candidate c at ((k NB + m) C + c) 8) with NB = n/2
// and the candidates per CTA as run-time values, so the partner loops are real loops and the chain bookkeeping a
// warp-uniform counter.  Blocks are walked in tiles of two.  Block b is paired with the SH blocks that follow it
// cyclically -- SH = (NB-1)/2 for odd NB; for even NB SH = NB/2 - 1 plus the "diameter" block b + NB/2 for the blocks
// of the lower half -- which covers every pair of blocks once.  In a tile {m, m+1}: step 0 pairs block m with block m+1
// (registers) and block m+1 with block m+SH+1; steps 1 .. SH-1 pair both with block m+1+s.  Every row meets one term per
// step, so one counter tells all of them when a reciprocal chain is full (six terms).
template <int D, bool MAXPRO>
struct RtTile {
    unsigned int base, BS, DS, WR;  // DS: dimension stride (all stored blocks); WR: NB * BS, the wrap of the cyclic pairing
    f2 eps2;
    __device__ __forceinline__ unsigned int block_off(unsigned int mm) const { const unsigned int o = mm * BS; return min(o, o - WR); }
    __device__ __forceinline__ void load_block(f2 (&b)[D], unsigned int off) const {
#pragma unroll
        for (int k = 0; k < D; ++k) b[k] = lds2(base + k * DS + off);
    }
    // MaxPro: q^2 + eps of two pairs; maximin: their squared distances
    __device__ __forceinline__ f2 term2(const f2 (&a)[D], const f2 (&b)[D]) const {
        if (MAXPRO) {
            f2 q = sub2(a[0], b[0]);
#pragma unroll
            for (int k = 1; k < D; ++k) q = mul2(q, sub2(a[k], b[k]));
            return fma2(q, q, eps2);
        }
        const f2 d0 = sub2(a[0], b[0]);
        f2 q = mul2(d0, d0);
#pragma unroll
        for (int k = 1; k < D; ++k) { const f2 dk = sub2(a[k], b[k]); q = fma2(dk, dk, q); }
        return q;
    }
    // one block's two rows as broadcast pairs, and the term of the pair inside the block
    __device__ __forceinline__ float rows_of(const f2 (&A)[D], f2 (&X)[2][D]) const {
        float q_in = 1.0f, s_in = 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float a0, a1;
            unpk(A[k], a0, a1);
            X[0][k] = pk(a0, a0); X[1][k] = pk(a1, a1);
            q_in = k == 0 ? a0 - a1 : q_in * (a0 - a1);
            s_in = fmaf(a0 - a1, a0 - a1, s_in);
        }
        float e0, e1;
        unpk(eps2, e0, e1);
        return MAXPRO ? fmaf(q_in, q_in, e0) : s_in;
    }
};

// running result of a candidate: MaxPro -- sum of reciprocals (chains + directly inverted terms); maximin -- minimum
template <bool MAXPRO>
struct RtAcc {
    f2 vsum;
    float direct, mn;
    __device__ __forceinline__ void reset() { vsum = pk(0.0f, 0.0f); direct = 0.0f; mn = CUDART_INF_F; }
    __device__ __forceinline__ void single(float t) { if (MAXPRO) direct += rcp_approx(t); else mn = fminf(mn, t); }
    __device__ __forceinline__ void single2(f2 t) { float lo, hi; unpk(t, lo, hi); if (MAXPRO) direct += rcp_approx(lo) + rcp_approx(hi); else mn = min3(mn, lo, hi); }
    __device__ __forceinline__ float result() const { float v0, v1; unpk(vsum, v0, v1); return MAXPRO ? v0 + v1 + direct : mn; }
};

template <int D, int GEN, bool MAXPRO, bool TRACK, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) search_screen_rt_kernel(ScreenParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long cta_thr[2];
    // Lanes per candidate G: 1 while 256 designs fit the shared memory of an SM (n d up to ~220), 2 while 128 do, else 4.
    // The lanes of a candidate sit in different parts of the warp (G = 2: lanes 0-15 / 16-31 of the same 16 candidates;
    // G = 4: quarter-warps of the same 8).  A 64-bit shared-memory access is served per half-warp, so with G = 2 the
    // two roles never collide; with G = 4 the two roles of a half-warp always read blocks two apart, and the block
    // stride carries 32 bytes of padding that puts those into the two halves of the banks.  Each lane builds every
    // G-th column and scores every G-th tile; the partial results meet in one or two shuffles.
    const int G = p.group;
    const int lane = threadIdx.x & 31, per_warp = 32 / G;
    const int c = (int)(threadIdx.x >> 5) * per_warp + lane % per_warp, g = lane / per_warp;
    // An odd number of points: the first n - 1 rows are paired block-wise as usual; the last row sits alone in one more
    // block (low half) and meets every block in a pass of its own.
    const bool n_odd = (p.n & 1) != 0;
    const int C = (int)blockDim.x / G, NB = p.n / 2, NBS = NB + (n_odd ? 1 : 0);  // paired blocks, stored blocks
    const unsigned int BS = (unsigned int)C * 8u + (G == 4 ? 32u : 0u), DS = (unsigned int)NBS * BS;
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(smem_raw) + (unsigned int)c * 8u;
    const bool odd = (NB & 1) != 0;
    const int SH = odd ? (NB - 1) / 2 : NB / 2 - 1;  // >= 1 (n >= 8)
    const unsigned int half = (unsigned int)(NB / 2);
    constexpr int kChainTerms = 6;

    const unsigned long long gtid = (unsigned long long)blockIdx.x * C + c;
    const unsigned long long gstep = (unsigned long long)gridDim.x * C;
    const unsigned long long total = p.hi - p.lo;
    const unsigned long long rounds = (total + gstep - 1) / gstep;
    if (threadIdx.x == 0) { cta_thr[0] = MAXPRO ? 0x7FF0000000000000ull : 0ull; cta_thr[1] = 0x7FF0000000000000ull; }
    __syncthreads();
    RtTile<D, MAXPRO> T{base, BS, DS, (unsigned int)NB * BS, pk(p.eps_s, p.eps_s)};

    for (unsigned long long rnd = 0; rnd < rounds; ++rnd) {
        const unsigned long long off = gtid + rnd * gstep;
        const bool active = off < total;
        const unsigned long long idx = p.lo + off;
        double s = 0.0;
        unsigned int gap_min = 0xFFFFFFFFu;
        if (active) {
            const unsigned long long stream = p.seed + idx;  // src/lib.rs:70
            constexpr bool kAllCols = MAXPRO && D > 5;  // the power-of-two scale of the MaxPro terms: in column 0, or spread over all
            if (G == 1) gap_min = screen_generate<D, GEN, kAllCols, TRACK>(base, NBS, BS, DS, (uint32_t)stream, (uint32_t)(stream >> 32), p.col0_scale, p.jbits, 0, 1, D, n_odd, p.col_scale);
            else if (G == 2) gap_min = screen_generate<(D + 1) / 2, GEN, kAllCols, TRACK>(base, NBS, BS, DS, (uint32_t)stream, (uint32_t)(stream >> 32), p.col0_scale, p.jbits, g, 2, D, n_odd, p.col_scale);
            else gap_min = screen_generate<(D + 3) / 4, GEN, kAllCols, TRACK>(base, NBS, BS, DS, (uint32_t)stream, (uint32_t)(stream >> 32), p.col0_scale, p.jbits, g, 4, D, n_odd, p.col_scale);
        }
        if (G > 1) __syncwarp();  // the design is complete before either lane scores it
        if (active) {
            double acc = 0.0;
            float mn_all = CUDART_INF_F;
#pragma unroll 1
            for (int m = 2 * g; m < NB; m += 2 * G) {
                RtAcc<MAXPRO> R;
                R.reset();
                if (m + 1 < NB) {
                    // ---- a tile of two blocks
                    f2 A0[D], A1[D], X[2][2][D];
                    T.load_block(A0, (unsigned int)m * BS);
                    T.load_block(A1, (unsigned int)(m + 1) * BS);
                    R.single(T.rows_of(A0, X[0]));
                    R.single(T.rows_of(A1, X[1]));
                    Chain2 ch[2][2];
                    int cnt = 0;
                    auto feed = [&](int b, const f2 (&P)[D]) {  // one term for each of the two rows of block b (cnt is warp-uniform)
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const f2 pr = T.term2(X[b][r], P);
                            if (!MAXPRO) { R.single2(pr); continue; }
                            if (cnt == 0) ch[b][r].start(pr); else ch[b][r].push(pr);
                        }
                    };
                    auto flush = [&]() {
                        if (MAXPRO) R.vsum = add2(R.vsum, add2(add2(ch[0][0].value(), ch[0][1].value()), add2(ch[1][0].value(), ch[1][1].value())));
                        cnt = 0;
                    };
                    {   // step 0
                        f2 E[D];
                        T.load_block(E, T.block_off((unsigned int)(m + 1 + SH)));
                        feed(0, A1);
                        feed(1, E);
                        cnt = 1;
                    }
#pragma unroll 1
                    for (int st = 1; st < SH; ++st) {
                        if (cnt == kChainTerms) flush();
                        f2 B[D];
                        T.load_block(B, T.block_off((unsigned int)(m + 1 + st)));
                        feed(0, B);
                        feed(1, B);
                        ++cnt;
                    }
                    flush();
                    if (!odd) {  // even NB: the diameter partner of the blocks of the lower half
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            if ((unsigned int)(m + b) < half) {
                                f2 Dm[D];
                                T.load_block(Dm, ((unsigned int)(m + b) + half) * BS);
                                R.single2(T.term2(X[b][0], Dm));
                                R.single2(T.term2(X[b][1], Dm));
                            }
                        }
                    }
                } else {
                    // ---- the odd block out (odd NB): block NB-1 against blocks 0 .. SH-1
                    f2 A0[D], X[2][D];
                    T.load_block(A0, (unsigned int)m * BS);
                    R.single(T.rows_of(A0, X));
                    Chain2 ch[2];
                    int cnt = 0;
#pragma unroll 1
                    for (int st = 0; st < SH; ++st) {
                        if (cnt == kChainTerms) {
                            if (MAXPRO) R.vsum = add2(R.vsum, add2(ch[0].value(), ch[1].value()));
                            cnt = 0;
                        }
                        f2 B[D];
                        T.load_block(B, (unsigned int)st * BS);
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            const f2 pr = T.term2(X[r], B);
                            if (!MAXPRO) { R.single2(pr); continue; }
                            if (cnt == 0) ch[r].start(pr); else ch[r].push(pr);
                        }
                        ++cnt;
                    }
                    if (MAXPRO) R.vsum = add2(R.vsum, add2(ch[0].value(), ch[1].value()));
                }
                if (MAXPRO) acc += (double)R.result(); else mn_all = fminf(mn_all, R.result());
            }
            if (n_odd) {
                // ---- the last row (odd n) against every block; with two lanes per candidate each takes every other block
                RtAcc<MAXPRO> R;
                R.reset();
                f2 AL[D], XL[D];
                T.load_block(AL, (unsigned int)NB * BS);
#pragma unroll
                for (int k = 0; k < D; ++k) { float a0, a1; unpk(AL[k], a0, a1); XL[k] = pk(a0, a0); }
                Chain2 chl;
                int cnt = 0;
#pragma unroll 1
                for (int m = g; m < NB; m += G) {
                    f2 B[D];
                    T.load_block(B, (unsigned int)m * BS);
                    const f2 pr = T.term2(XL, B);
                    if (!MAXPRO) { R.single2(pr); continue; }
                    if (cnt == kChainTerms) { R.vsum = add2(R.vsum, chl.value()); cnt = 0; }
                    if (cnt == 0) chl.start(pr); else chl.push(pr);
                    ++cnt;
                }
                if (MAXPRO && cnt > 0) R.vsum = add2(R.vsum, chl.value());
                if (MAXPRO) acc += (double)R.result(); else mn_all = fminf(mn_all, R.result());
            }
            s = MAXPRO ? acc : (double)mn_all;
        }
        // the partial results of a candidate's lanes (role 0 ends up with the whole); also: every lane is done reading
        // before the design is overwritten
        for (int w = 16; w >= 32 / G && G > 1; w >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, s, w);
            s = MAXPRO ? s + other : fmin(s, other);
            if (TRACK) gap_min = min(gap_min, __shfl_xor_sync(0xffffffffu, gap_min, w));
        }
        if (active && g == 0 && p.debug_scores) p.debug_scores[off] = s;
        screen_survivors<MAXPRO>(p, cta_thr, rnd == 0, active && g == 0, s, idx, gap_min);
    }
}

// candidates per CTA of the general kernel for a shape (0: the shape is not served) and the lanes per candidate
static int screen_rt_cands(unsigned long long n, unsigned long long d, int* group) {
    if (group) *group = 1;
    if (n < 8 || n > 256 || d < 1 || d > 10) return 0;
    const unsigned long long rows = n + (n & 1);  // an odd design stores one more (unused) row
    size_t c = (kSmemBudget - kSmemSlack) / (rows * d * sizeof(float));
    if (c >= 256 && d <= 8) return (int)((c > 512 ? 512 : c) / 32 * 32);
    if (c >= 128 && d <= 8) {  // two lanes per candidate: 16 candidates per warp
        if (group) *group = 2;
        return (int)(c / 16 * 16);
    }
    // four lanes per candidate, 8 candidates per warp, 32 bytes of padding per row block; d = 9, 10: at most 256 threads
    // (the kernel needs more than 128 registers there)
    const size_t per_block_row = (kSmemBudget - kSmemSlack) / (d * (rows / 2));  // bytes one row block of all candidates may take
    if (per_block_row < 32 + 8 * 32) return 0;
    c = (per_block_row - 32) / 8 / 8 * 8;
    if (c > 120) c = 120;
    if (d > 8 && c > 64) c = 64;
    if (group) *group = 4;
    return c >= 32 ? (int)c : 0;
}

template <int D>
cudaError_t launch_screen_rt_d(const SearchLaunch& L, const ScreenParams& p, int cands) {
    const size_t nbs = (size_t)(L.n + (L.n & 1)) / 2;
    const size_t smem = (size_t)D * nbs * ((size_t)cands * 8 + (p.group == 4 ? 32 : 0));
    const int threads = cands * p.group;
    constexpr int kMaxT = D <= 8 ? 512 : 256;
    if (threads > kMaxT) return cudaErrorInvalidValue;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<L.grid, threads, smem, L.stream>>>(p);
        return cudaGetLastError();
    };
    if (L.metric == 0) {
        // the second (adjacent-gap) bound costs ~2.5 instructions per generated element: tracked where the first one is too
        // wide to keep the survivors few -- d >= 4, and d = 3 beyond n = 160 (measured: tools/screen_sweep.py)
        constexpr bool kAlways = D >= 4, kNever = D <= 2;
        const bool track = kAlways || (!kNever && L.n > 160);
        if constexpr (kAlways) return p.gen == 1 ? go(search_screen_rt_kernel<D, 1, true, true, kMaxT>) : go(search_screen_rt_kernel<D, 2, true, true, kMaxT>);
        else if constexpr (kNever) return p.gen == 1 ? go(search_screen_rt_kernel<D, 1, true, false, kMaxT>) : go(search_screen_rt_kernel<D, 2, true, false, kMaxT>);
        else if (track) return p.gen == 1 ? go(search_screen_rt_kernel<D, 1, true, true, kMaxT>) : go(search_screen_rt_kernel<D, 2, true, true, kMaxT>);
        else return p.gen == 1 ? go(search_screen_rt_kernel<D, 1, true, false, kMaxT>) : go(search_screen_rt_kernel<D, 2, true, false, kMaxT>);
    }
    return p.gen == 1 ? go(search_screen_rt_kernel<D, 1, false, false, kMaxT>) : go(search_screen_rt_kernel<D, 2, false, false, kMaxT>);
}

template <int N, int D, int C>
cudaError_t launch_screen_nd(const SearchLaunch& L, const ScreenParams& p) {
    constexpr size_t smem = (size_t)N * D * C * sizeof(float);
    static_assert(smem <= kSmemBudget, "candidates per CTA do not fit shared memory");
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<L.grid, C, smem, L.stream>>>(p);
        return cudaGetLastError();
    };
    if (L.metric == 0) return p.gen == 1 ? go(search_screen_kernel<N, D, C, 1, true>) : go(search_screen_kernel<N, D, C, 2, true>);
    return p.gen == 1 ? go(search_screen_kernel<N, D, C, 1, false>) : go(search_screen_kernel<N, D, C, 2, false>);
}

}  // namespace

// Shapes with a screening kernel; *cands = candidates per CTA, *threads = its block size.  The specialised kernel serves
// n = 50, d = 2 or 3; the general one any n from 8 to 256 whose design fits shared memory 32 times (n d <= ~1700, d <= 10),
// for maximin (2-3.6x the fp64 kernels on every shape measured) and, from d = 2 up, for MaxPro (1.15-2.1x:
// tools/screen_sweep.py, profiles/r2w_screen_*.txt; at d = 1 the fp64 kernel is as fast).  With the first bound alone
// MaxPro stopped paying at d = 5 and at (100,4), (200,3): the survivors swamped stage 2 (0.5-0.7x, profiles/r2q_*, r2t_*);
// the adjacent-gap bound (header) keeps them to a fraction of a percent at every d.
bool search_screen_supported(unsigned long long n, unsigned long long d, int metric, int* cands, int* threads) {
    int c = 0, group = 1;
    if (metric == 0 || metric == 1) {
        if (n == 50 && d == 3) c = 384;
        else if (n == 50 && d == 2) c = 512;
        else if (!cfg("MAXPRO_NO_SCREEN_RT") && (metric == 1 || d >= 2)) c = screen_rt_cands(n, d, &group);
    }
    if (cands) *cands = c;
    if (threads) *threads = c * group;
    return c != 0;
}

// Host: the constants of the bound for a shape (see the header comment).
// jitter bits stage 1 keeps for a design of n points: the most that leaves y = i 2^jbits + ... below 2^24
static int screen_jbits(unsigned long long n) {
    int lg = 0;
    while ((1ull << lg) < n) ++lg;
    const int jb = 24 - lg;
    return jb > 18 ? 18 : jb;  // 18 is what the specialised kernel is built for; more buys nothing
}

static ScreenBound screen_bound(unsigned long long n, unsigned long long d, float* col0_scale, float* eps_s, int* scale_exp = nullptr) {
    const int jb = screen_jbits(n);
    const double unit = ldexp((double)n, jb);  // y = x * n * 2^jbits
    // exact power of two on column 0 that puts G^2 eps next to 2^-20
    const double log2_g2eps = log2(kEps) + 2.0 * (double)d * log2(unit);
    const int a = (int)llround(0.5 * (log2_g2eps + 20.0));
    const double G = pow(unit, (double)d) * ldexp(1.0, -a);
    *col0_scale = (float)ldexp(1.0, -a);  // d >= 6: below the fp32 range; the callers spread 2^-a over the columns (scale_exp)
    if (scale_exp) *scale_exp = a;
    *eps_s = (float)(kEps * G * G);
    ScreenBound b;
    b.g2 = G * G;
    b.eps = (double)*eps_s / b.g2;
    const double eta = ldexp(1.0, -jb) / (double)n;
    b.dq = (double)d * eta * pow(1.0 + eta, (double)d - 1.0);
    // gmin = min over q >= 0 of (q^2 + eps) / ((q + dq)^2 + eps), scanned on a log grid around sqrt(eps), 2 % margin
    double g = 1.0;
    for (int t = -4000; t <= 4000; ++t) {
        const double q = sqrt(b.eps) * exp2((double)t / 200.0);
        const double r = (q * q + b.eps) / ((q + b.dq) * (q + b.dq) + b.eps);
        if (r < g) g = r;
    }
    b.gmin = g * 0.98;
    b.arith = ldexp(1.0, d <= 3 ? -18 : -17);  // (2d + 25) 2^-24 at most: products, fma, six chain pushes, reciprocal, sums
    b.tau = (4.0 * (double)d + 4.0) * ldexp(1.0, -52);
    // maximin (see the header): units of u = n 2^jbits, no scaling
    b.mm_c1 = 2.0 * sqrt((double)d) * (1.0 + ldexp(1.0, -20));
    b.mm_c0 = (double)d;
    b.mm_arith = ((double)d + 1.0) * ldexp(1.0, -23);
    return b;
}

// list: survivor indices (cap entries), count: number appended (zeroed by the caller's init kernel)
cudaError_t launch_search_screen(const SearchLaunch& L, unsigned long long* list, unsigned int* count, unsigned int cap,
                                 double* debug_scores) {
    ScreenParams p;
    p.gen = stream_version();
    p.n = (int)L.n; p.d = (int)L.d;
    p.group = 1;
    p.jbits = screen_jbits(L.n);
    p.seed = L.seed; p.lo = L.lo; p.hi = L.hi;
    int scale_exp = 0;
    p.bound = screen_bound(L.n, L.d, &p.col0_scale, &p.eps_s, &scale_exp);
    for (int k = 0; k < 10; ++k) {  // 2^-a as d factors whose exponents differ by at most one
        const int dd = (int)L.d, share = scale_exp / dd + (k < scale_exp % dd ? 1 : 0);
        p.col_scale[k] = k < dd ? (float)ldexp(1.0, -share) : 1.0f;
    }
    if (L.metric != 0) p.col0_scale = 1.0f;  // maximin works in plain units
    p.list = list; p.count = count; p.cap = cap;
    p.debug_scores = debug_scores;
    if (L.n == 50 && L.d == 3) return launch_screen_nd<50, 3, 384>(L, p);
    if (L.n == 50 && L.d == 2) return launch_screen_nd<50, 2, 512>(L, p);
    const int cands = screen_rt_cands(L.n, L.d, &p.group);
    if (cands == 0 || cands * p.group != L.threads) return cudaErrorInvalidValue;
    switch (L.d) {
        case 1: return launch_screen_rt_d<1>(L, p, cands);
        case 2: return launch_screen_rt_d<2>(L, p, cands);
        case 3: return launch_screen_rt_d<3>(L, p, cands);
        case 4: return launch_screen_rt_d<4>(L, p, cands);
        case 5: return launch_screen_rt_d<5>(L, p, cands);
        case 6: return launch_screen_rt_d<6>(L, p, cands);
        case 7: return launch_screen_rt_d<7>(L, p, cands);
        case 8: return launch_screen_rt_d<8>(L, p, cands);
        case 9: return launch_screen_rt_d<9>(L, p, cands);
        case 10: return launch_screen_rt_d<10>(L, p, cands);
    }
    return cudaErrorInvalidValue;
}

// tests: S_true of a scaled score, and the survivor threshold a scaled score implies, in true units
void screen_debug_constants(unsigned long long n, unsigned long long d, double* g2, double* threshold_ratio_at, double s_true) {
    float c0, es;
    const ScreenBound b = screen_bound(n, d, &c0, &es);
    if (g2) *g2 = b.g2;
    if (threshold_ratio_at) *threshold_ratio_at = screen_threshold(s_true / b.g2, b) * b.g2 / s_true;
}

}  // namespace mp

// Shared device/host helpers for the maxpro B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mp {

constexpr double kEps = 1e-12;  // src/maxpro_utils.rs:35
// Terms a RecipChain may hold before a flush.  The chain's denominator starts at 2^1000 instead of 1
// (num/den is scale-free), so it has the whole exponent range below it: 2^1000 * (1e-12)^k >= 1e-300
// holds for k <= 50, and num <= k * 2^1000.
constexpr int kMaxChainTerms = 48;

// Winner record that leaves the chip: 16 bytes.
struct __align__(16) BestRec {
    double score;
    unsigned long long index;
};

constexpr unsigned long long kNoIndex = 0xFFFFFFFFFFFFFFFFull;

// Fold of src/lib.rs:85-97 with `a` the earlier (left) element: the left one
// survives only on a strict win, so equal scores keep the LATER index.  Records
// carry global indices, so "later" is simply the larger index and the fold is
// order-free.  kNoIndex marks the identity (lib.rs:77-83).
template <bool MINIMIZE>
__host__ __device__ __forceinline__ bool better(double sa, unsigned long long ia,
                                                double sb, unsigned long long ib) {
    if (ib == kNoIndex) return true;
    if (ia == kNoIndex) return false;
    if (sa == sb) return ia > ib;
    return MINIMIZE ? (sa < sb) : (sa > sb);
}

template <bool MINIMIZE>
__device__ __forceinline__ void warp_reduce_best(double& s, unsigned long long& i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double so = __shfl_xor_sync(0xffffffffu, s, off);
        unsigned long long io = __shfl_xor_sync(0xffffffffu, i, off);
        if (better<MINIMIZE>(so, io, s, i)) { s = so; i = io; }
    }
}

__device__ __forceinline__ double warp_reduce_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Warp minimum of values that are >= +0 (squared distances, every caller's case): the bit pattern orders like
// the number, so two REDUX.MIN (high word, then low word among the lanes that hold the high minimum) replace
// five shuffle levels of fmin.  A NaN pattern sorts above +inf and is returned only if every lane holds one,
// which is what the fmin tree gives.
__device__ __forceinline__ double warp_reduce_min(double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned int hi = (unsigned int)(bits >> 32), lo = (unsigned int)bits;
    const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xFFFFFFFFu);
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

// Minimum of two squared distances on their bit patterns.  For values >= +0 the unsigned integer
// order is the numeric order; a NaN pattern sorts above +inf and is never taken, which is what the
// reference does with it (`dist < min_dist` is false, src/maximin_utils.rs:61).  The comparison then
// runs on the integer pipes: fmin(double) costs a DSETP on the FP64 pipe plus four select/logic
// instructions, and the FP64 pipe is what the maximin kernels are short of.
__device__ __forceinline__ double min_sqdist(double a, double b) {
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
    return __longlong_as_double((long long)(ub < ua ? ub : ua));
}

// 1/x for normal, positive x: MUFU.RCP64H seed (>= 20 good bits) + two Newton
// steps (20 -> 40 -> 80 bits).  Relative error ~1 ulp; no special-case branch.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Running sum of reciprocals without a divide per term:
//   num/den + 1/p = (num*p + den) / (den*p)
// Two FP64 ops per pushed term; one reciprocal per flush.  Safe while the product of pushed terms
// stays in range: every p is in [1e-12, ~1] for designs in the unit cube, so kMaxChainTerms terms
// keep den inside [1e-300, 2^1000] (see above).
struct RecipChain {
    double num, den;
    __device__ __forceinline__ void reset() { num = 0.0; den = 0x1p1000; }
    __device__ __forceinline__ void push(double p) { num = fma(num, p, den); den *= p; }
    // acc + num/den with the quotient good to ~2^-40 (one Newton step on the 20-bit MUFU seed,
    // the residual folded into the final FMA): every partial sum is then within 1e-12 relative,
    // three orders inside the 1e-9 parity bar, for 4 FP64 instructions.  den is a normal number
    // throughout, so are the seed and the residual.
    __device__ __forceinline__ double flush_into(double acc) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
        const double e = fma(-den, r, 1.0);
        const double nr = num * r;
        const double v = fma(nr, e, nr);  // num * r * (1 + e)
        reset();
        return acc + v;
    }
};

}  // namespace mp

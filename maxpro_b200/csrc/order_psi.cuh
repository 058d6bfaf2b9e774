// Near-tie resolution of the greedy ordering (src/order.rs:84-99) on the criterion itself.
//
// order_design compares psi(prefix + candidate) = ((S_prefix + acc_c) / pairs)^(1/d) between pool points with a
// strict compare, so equal psi -> first pool position.  psi is monotone in the running value acc_c, and the
// kernels find the best acc_c with integer-key reductions; but two acc a few ulp apart can round to ONE psi
// (the 1/d-th power contracts relative differences by d), and then the reference takes the earlier position
// where argmin-acc takes the smaller acc.  So after the argbest on acc every kernel looks for other pool points
// inside a band around the best acc in which psi can still coincide, and if there is one it evaluates psi for
// the contenders exactly as the reference does -- same S_prefix (summed in pick order), same pair count, same
// division, and pow rounded like libm's (pow_glibc.cuh) -- and applies the position rule to THAT.
#pragma once
#include "pow_glibc.cuh"

namespace mp {

// n(n-1)/2 of the prefix plus one candidate (src/maxpro_utils.rs:81), m1 = points in it
__device__ __forceinline__ double order_pairs(unsigned int m1) {
    const double n = (double)m1;
    return __ddiv_rn(__dmul_rn(n, __dsub_rn(n, 1.0)), 2.0);
}

__device__ __forceinline__ double order_psi(double s_prefix, double acc, double n_pairs, double inv_d) {
    return pow_glibc(__ddiv_rn(__dadd_rn(s_prefix, acc), n_pairs), inv_d);  // src/maxpro_utils.rs:82
}

// Half-width of the band: psi((S + a)(1 + t)) ~ psi(S + a)(1 + t/d), so beyond t = (4d + 4) 2^-52 the two criteria
// are at least two ulp apart -- more than the rounding of pow (< 1 ulp) can close.
__device__ __forceinline__ double order_tie_band(double s_prefix, double acc_best, unsigned int d) {
    return __dmul_rn(__dadd_rn(s_prefix, acc_best), (4.0 * (double)d + 4.0) * 0x1p-52);
}

// High word of the integer key of a non-negative running value (smaller key = better; ~bits when the larger value
// wins).  The kernels carry the SECOND-best key of a thread / warp / CTA in this form to detect contenders: dropping
// the low word rounds the key towards "better", so `second_hi <= order_key_hi(band limit)` never misses one (it can
// raise a false alarm when two values agree to 2^-20, which only costs the slow path).
constexpr unsigned int kNoKeyHi = 0xFFFFFFFFu;
template <bool MINIMIZE>
__device__ __forceinline__ unsigned int order_key_hi(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    if (v != v) return kNoKeyHi;
    return (unsigned int)((MINIMIZE ? b : ~b) >> 32);
}

}  // namespace mp

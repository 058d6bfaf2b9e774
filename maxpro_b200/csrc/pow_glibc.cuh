// pow(x, y) with the arithmetic of glibc >= 2.28 (ARM Optimized Routines pow, FMA code path), host and device.
//
// Why the engine carries its own pow.  The reference compares criteria AFTER the 1/d-th power:
// maxpro_criterion ends in `(S / n_pairs).powf(1.0 / d)` (src/maxpro_utils.rs:81-82), build_lhd folds on
// that value (src/lib.rs:85-97: equal criteria -> later index) and order_design compares it between pool
// points (src/order.rs:84-99: equal criteria -> first pool position).  Two raw sums a few ulp apart can
// collapse into one criterion value, and then the tie rule -- not the sum -- picks the design.  Which sums
// collapse depends on the rounding of pow to the last bit; Rust's f64::powf is the platform libm's pow,
// and CUDA's pow is a different function (<= 2 ulp).  So the kernels break near-ties with THIS function:
// the same tables (pow_glibc_tables.inc, read out of libm.so.6 by tools/extract_pow_tables.py), the same
// polynomials and the same operation order as the __pow_fma variant glibc selects on any x86-64 host
// with FMA, transcribed from its instruction sequence.  tests/test_pow_glibc.py builds this header for
// the host and compares it with libm bit for bit (4e8 arguments checked when it was written: 0 differences).
//
// Exact-path domain: x positive and finite, 2^-65 <= |y| < 2^63, |y log x| < 1024 with a normal (or infinite)
// result.  That covers every call the engine makes (x = S / n_pairs > 0, y = 1/d) with a margin of hundreds of
// binades.  Anything else (x <= 0, NaN, subnormal results, huge y) goes to the toolkit's pow (device) / std::pow (host).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif

namespace mp {

#if defined(__CUDA_ARCH__)
#define MP_POW_TABLE(name, count) static __device__ const unsigned long long name[count]
#else
#define MP_POW_TABLE(name, count) static const unsigned long long name[count]
#endif
#include "pow_glibc_tables.inc"
#undef MP_POW_TABLE

namespace powdetail {

MP_HD double as_double(unsigned long long u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
MP_HD unsigned long long as_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
// every product and sum below is a separately rounded IEEE operation unless written as fma: the intrinsics
// keep nvcc from contracting what glibc's build did not contract (and g++ gets -ffp-contract=off)
MP_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
MP_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
MP_HD double sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
MP_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

}  // namespace powdetail

MP_HD double pow_glibc(double x, double y) {
    using namespace powdetail;
    unsigned long long ix = as_bits(x);
    const unsigned long long iy = as_bits(y);
    const unsigned int topx = (unsigned int)(ix >> 52), topy = (unsigned int)(iy >> 52) & 0x7ffu;
    // x: positive, finite, non-zero; subnormal x is normalised as in glibc.  y: 2^-65 <= |y| < 2^63.
    bool exact = topx < 0x7ffu && ix != 0ull && (topy - 0x3beu) <= 0x7fu;
    if (exact && topx == 0u) {
        ix = as_bits(mul(x, 0x1p52));
        ix -= 52ull << 52;
    }
    if (exact) {
        // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1) as hi + lo
        const unsigned long long tmp = ix - 0x3fe6955500000000ull;
        const unsigned int i = (unsigned int)(tmp >> 45) & 127u;
        const long long k = (long long)tmp >> 52;
        const double z = as_double(ix - (tmp & (0xfffull << 52)));
        const double kd = (double)k;
        const double invc = as_double(kPowLogTab[3 * i]), logc = as_double(kPowLogTab[3 * i + 1]), logctail = as_double(kPowLogTab[3 * i + 2]);
        const double A0 = -0x1p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1, A3 = 0x1.999999959554ep-1,
                     A4 = -0x1.555555529a47ap-1, A5 = -0x1.2495b9b4845e9p+0, A6 = 0x1.0002b8b263fc3p+0;
        const double t1 = fma_(kd, 0x1.62e42fefa3800p-1, logc);
        const double lo1 = fma_(kd, 0x1.ef35793c76730p-45, logctail);
        const double r = fma_(z, invc, -1.0);
        const double ar = mul(r, A0);
        const double q12 = fma_(r, A2, A1);
        const double q34 = fma_(r, A4, A3);
        const double t2 = add(r, t1);
        const double lo2 = add(sub(t1, t2), r);
        const double ar2 = mul(r, ar);
        const double ar3 = mul(r, ar2);
        const double lo3 = fma_(ar, r, -ar2);
        const double hi0 = add(t2, ar2);
        const double q56 = fma_(r, A6, A5);
        const double lo4 = add(sub(t2, hi0), ar2);
        double q = fma_(q56, ar2, q34);
        q = fma_(ar2, q, q12);
        double lo = add(lo1, lo2);
        lo = add(lo, lo3);
        lo = add(lo, lo4);
        lo = fma_(ar3, q, lo);
        const double hi = add(hi0, lo);
        const double tail = add(sub(hi0, hi), lo);
        // ---- y * log(x) as ehi + elo
        const double ehi = mul(y, hi);
        const double elo = fma_(y, tail, fma_(hi, y, -ehi));
        // ---- exp_inline
        const unsigned int abstop = (unsigned int)(as_bits(ehi) >> 52) & 0x7ffu;
        if (abstop - 0x3c9u <= 0x3fu) {  // 2^-54 <= |y log x| < 1024
            double kd2 = fma_(ehi, 0x1.71547652b82fep+7, 0x1.8p52);
            const unsigned long long ki = as_bits(kd2);
            kd2 = sub(kd2, 0x1.8p52);
            double rr = fma_(kd2, -0x1.62e42fefa0000p-8, ehi);
            rr = fma_(kd2, -0x1.cf79abc9e3b3ap-47, rr);
            rr = add(elo, rr);
            const unsigned int idx = 2u * ((unsigned int)ki & 127u);
            const double tl = as_double(kPowExpTab[idx]);
            const unsigned long long sbits = kPowExpTab[idx + 1] + (ki << 45);
            const double p23 = fma_(rr, 0x1.555555555543cp-3, 0x1.ffffffffffdbdp-2);
            const double s0 = add(rr, tl);
            const double r2 = mul(rr, rr);
            const double p45 = fma_(rr, 0x1.1111167a4d017p-7, 0x1.55555cf172b91p-5);
            double t = fma_(p23, r2, s0);
            const double r4 = mul(r2, r2);
            t = fma_(p45, r4, t);
            if (abstop != 0x408u) return fma_(t, as_double(sbits), as_double(sbits));
            // 512 <= |y log x| < 1024 (glibc's specialcase): the exponent of `scale` may have left the range, so the
            // result is formed with a rescaled scale and brought back by an exact power of two
            if ((ki & 0x80000000ull) == 0) {  // k > 0: overflows to +inf exactly where glibc's does
                const double sc = as_double(sbits - (1009ull << 52));
                return mul(0x1p1009, fma_(t, sc, sc));
            }
            const double sc = as_double(sbits + (1022ull << 52));
            const double yv = add(sc, mul(sc, t));  // not fused in glibc's build: the product is shared with the subnormal branch
            if (yv >= 1.0) return mul(0x1p-1022, yv);  // normal result; below that glibc rounds in the subnormal range
        } else if (abstop < 0x3c9u) {
            return add(1.0, ehi);  // |y log x| < 2^-54
        } else if (abstop < 0x7ffu) {
            return (as_bits(ehi) >> 63) ? 0.0 : as_double(0x7ff0000000000000ull);  // |y log x| >= 1024: underflow / overflow
        }
    }
#ifdef MP_POW_NO_FALLBACK
    return -1.0;  // host test build: shows which arguments leave the exact path
#else
    return pow(x, y);
#endif
}

}  // namespace mp

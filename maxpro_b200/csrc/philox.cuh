// Philox4x32-10 and the "LHD-Philox v1" stream layout (DESIGN.md, section RNG).
// The reference draws from rand 0.9.2 StdRng (src/lhd.rs:119,126, src/anneal.rs:29-31);
// that stream is not pinned by any reference test, so the engine defines its own
// counter-based stream: a draw is a pure function of (stream id, purpose, position),
// which is what lets a kernel regenerate any candidate from its index alone.
#pragma once
#include <stdint.h>

namespace mp {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// generate_lhd stream: FIXED key, stream id in the counter -> the ten round keys
// are compile-time constants and cost no instructions in the search kernel.
//   Philox(ctr = {b, k, stream_lo, stream_hi}, key = {kKeyLhd0, kKeyLhd1})
//   -> {jit, idx} of element 2b and {jit, idx} of element 2b+1 of column k.
constexpr uint32_t kKeyLhd0 = 0x4C484431u;  // "LHD1"
constexpr uint32_t kKeyLhd1 = 0x243F6A88u;
// annealer streams: key = {stream_lo, stream_hi}, ctr = {t_lo, t_hi, block, domain}
constexpr uint32_t kDomainSwap = 0x53574150u;  // "SWAP": swap_rows draws + accept draw
constexpr uint32_t kDomainJacc = 0x4A414343u;  // "JACC": jitter accept draw
constexpr uint32_t kDomainJit = 0x4A495454u;   // "JITT": jitter deltas

struct Philox4 {
    uint32_t x, y, z, w;
};

// 32x32 -> 64 multiply returned as (hi, lo).  On the device a single
// mul.wide.u32 (one IMAD.WIDE.U32) written in PTX: the C++ form makes nvcc add a
// zero carry to every high word.
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%1, %0}, t;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
#else
    uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(kPhiloxM0, c0, hi0, lo0);
        mulhilo32(kPhiloxM1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        c0 = n0;
        c2 = n2;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return Philox4{c0, c1, c2, c3};
}

// v_i = (i + (jit + 1/2) 2^-32) / n = fma((double)(i*2^32 + jit), c1, c0): the stratum value of
// the LHD-Philox stream.  The 64-bit integer is assembled by register pairing (mov.b64), not by
// shift/or arithmetic, which nvcc otherwise expands into carry chains.
__device__ __forceinline__ double stratum_value(uint32_t i, uint32_t jit, double c1, double c0) {
    unsigned long long m;
    asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "r"(jit), "r"(i));
    return fma(__ull2double_rn(m), c1, c0);
}

// ---- LHD-mix v2: the design stream with the generator off the FP64 issue path ---------------------------
// v1 spends one Philox4x32-10 call (20 wide multiplies, 20 three-input logic ops) on every TWO elements, and in
// the fused search kernels those instructions do not hide behind the FP64 stream: they were 15 % of the headline
// kernel's time (profiles/README.md, r1g).  v2 keeps Philox where it buys independence between streams -- ONE call
// per column, keyed by (stream id, column) -- and draws the elements of a column from a counter-based 32-bit
// mixer under that key:
//     {A, B} = Philox4x32-10(ctr = {k, "LHD2", s_lo, s_hi}, key = {kKeyLhd0, kKeyLhd1}).{x, y | 1}
//     w_i    = mix32(A + i * B)            mix32: x ^= x>>16; x *= 0x735a2d97
//     P      = w_i * (i + 1)               (64 bits)
//     j_i    = P >> 32                     the Fisher-Yates index in [0, i], bias (i+1) 2^-32
//     jit_i  = P mod 2^32                  the fractional part of w_i (i+1) 2^-32: uniform and independent of j_i
//     v_i    = (i + (jit_i + 1/2) 2^-32) / n,  evaluated as  fma(2^52 + (i 2^32 + jit_i), c1, c0'),
//              c1 = 2^-32 / n,  c0' = c1/2 - 2^52 c1:  the integer is written into the mantissa of 2^52 (one
//              integer op instead of a 64-bit int -> double conversion); floor(v_i n) == i for every draw, n <= 2^18.
// mix32 is ONE xorshift-multiply round.  Its input is not a counter but a Weyl sequence A + i B whose start and odd step
// are both full-entropy Philox words of the column, so the sequence is equidistributed to begin with; what the round has
// to do is break its linearity (w_i, w_{i+1}, w_{i+2} collinear) in the HIGH bits, which are the ones the shuffle index and
// the leading jitter bits are made of.  A first multiply would be wasted (it maps the sequence onto another Weyl sequence
// with equally random start and step), and a second round moved nothing in the tests: tools/mixer_tests.py compares this
// round with the two-round finaliser (C. Wellons' hash-prospector constants) it replaced and with three other candidates
// on 7.5e7-sample contingency tables of (w_i, w_{i+lag}) bit fields, triples, Fisher-Yates index pairs, and on the
// permutation statistics of 40,000 designs -- no difference -- while a rotate-multiply mixer fails them at once.  In the
// fused kernels every generator instruction costs an issue cycle that the criterion does not get:
// the two-round finaliser was 5.6 ms of the screening kernel's 54.7 ms per 2^27 candidates.  Per element: 1 multiply,
// one wide multiply and 3 shift/logic ops, against 10 wide multiplies and 10 logic ops in v1.  tests/test_oracle.py holds
// the statistics (uniformity of every position, of adjacent orderings and differences, independence of neighbouring
// streams and columns, whole permutations for small n, KS and serial correlation of the jitter).
constexpr uint32_t kLhd2Tag = 0x4C484432u;  // "LHD2"

__host__ __device__ __forceinline__ uint32_t lhd2_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x735a2d97u;
    return x;
}

// one column's key: a = start, b = odd step of the Weyl sequence fed to the mixer
__host__ __device__ __forceinline__ void lhd2_column_key(uint32_t k, uint32_t s_lo, uint32_t s_hi, uint32_t& a, uint32_t& b) {
    const Philox4 w = philox4x32_10(k, kLhd2Tag, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
    a = w.x;
    b = w.y | 1u;
}

// element i of the column: Fisher-Yates index j in [0, i] and jitter word
__host__ __device__ __forceinline__ void lhd2_element(uint32_t a, uint32_t b, uint32_t i, uint32_t& j, uint32_t& jit) {
    mulhilo32(lhd2_mix32(a + i * b), i + 1u, j, jit);
}

// host: the additive constant of the v2 value formula (see above); c1 = ldexp(1.0 / n, -32)
__host__ __device__ __forceinline__ double lhd2_c0(double c1) { return 0.5 * c1 - 0x1p52 * c1; }

#ifdef __CUDACC__
__device__ __forceinline__ double lhd2_value(uint32_t i, uint32_t jit, double c1, double c0p) {
    return fma(__hiloint2double((int)(0x43300000u | i), (int)jit), c1, c0p);
}

// ---- version-independent front end used by every generator in the engine ---------------------------------------
// gen = 1: LHD-Philox v1, gen = 2: LHD-mix v2.  A draw block is the element pair (2b, 2b+1) of column k.
struct LhdBlock { uint32_t j0, jit0, j1, jit1; };
struct LhdKeyCache { uint32_t k, a, b; };  // v2: the key of the column a thread drew from last (k = ~0: none yet)

__device__ __forceinline__ LhdBlock lhd_draw_block(int gen, uint32_t b, uint32_t k, uint32_t s_lo, uint32_t s_hi, LhdKeyCache& kc) {
    LhdBlock r;
    if (gen == 1) {
        const Philox4 w = philox4x32_10(b, k, s_lo, s_hi, kKeyLhd0, kKeyLhd1);
        r.jit0 = w.x; r.j0 = __umulhi(w.y, 2u * b + 1u);
        r.jit1 = w.z; r.j1 = __umulhi(w.w, 2u * b + 2u);
    } else {
        if (kc.k != k) { kc.k = k; lhd2_column_key(k, s_lo, s_hi, kc.a, kc.b); }
        lhd2_element(kc.a, kc.b, 2u * b, r.j0, r.jit0);
        lhd2_element(kc.a, kc.b, 2u * b + 1u, r.j1, r.jit1);
    }
    return r;
}
#endif

#ifdef __CUDACC__
// v_i of either stream; c0 is the version's additive constant (host: lhd_c0)
__device__ __forceinline__ double lhd_value(int gen, uint32_t i, uint32_t jit, double c1, double c0) {
    return gen == 1 ? stratum_value(i, jit, c1, c0) : lhd2_value(i, jit, c1, c0);
}
#endif
inline double lhd_c0(int gen, double c1) { return gen == 1 ? 0.5 * c1 : lhd2_c0(c1); }

// The sequential half of the inside-out Fisher-Yates shuffle for one column whose draws were staged
// beforehand: a[i] holds the stratum value v_i and js[i] the index drawn for step i, and step i does
//   t = a[j_i];  a[i] = t;  a[j_i] = v_i.
// Steps before i touch positions below i only, so a[i] and js[i] of the next eight steps are read
// ahead of the dependent part: one shared-memory round trip per step instead of two.
__device__ __forceinline__ void fy_walk_column(double* a, const unsigned short* js, int n) {
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        int j[8];
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { j[u] = js[i + u]; v[u] = a[i + u]; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double t = a[j[u]];
            a[i + u] = t;
            a[j[u]] = v[u];
        }
    }
    for (; i < n; ++i) {
        const int j = js[i];
        const double v = a[i];
        const double t = a[j];
        a[i] = t;
        a[j] = v;
    }
}

#ifdef __CUDACC__
// The same walk for the v2 stream with nothing staged: the walking lane draws index and value of the next eight steps
// itself (lhd2_element + lhd2_value, independent of the chain) and then does the eight dependent steps.
__device__ __forceinline__ void fy_walk_column_v2(double* a, uint32_t ka, uint32_t kb, int n, double c1, double c0p) {
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        uint32_t j[8];
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t jit;
            lhd2_element(ka, kb, (uint32_t)(i + u), j[u], jit);
            v[u] = lhd2_value((uint32_t)(i + u), jit, c1, c0p);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double t = a[j[u]];  // j <= i + u: when j == i + u the value read is unspecified and overwritten below
            a[i + u] = t;
            a[j[u]] = v[u];
        }
    }
    for (; i < n; ++i) {
        uint32_t j, jit;
        lhd2_element(ka, kb, (uint32_t)i, j, jit);
        const double v = lhd2_value((uint32_t)i, jit, c1, c0p);
        const double t = a[j];
        a[i] = t;
        a[j] = v;
    }
}
#endif

// (w * bound) >> 32 : multiply-shift map of a 32-bit word onto [0, bound).
__host__ __device__ __forceinline__ uint32_t mulhi_bound(uint32_t w, uint32_t bound) {
    return (uint32_t)(((uint64_t)w * (uint64_t)bound) >> 32);
}

}  // namespace mp

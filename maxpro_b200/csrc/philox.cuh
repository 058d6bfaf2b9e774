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

// (w * bound) >> 32 : multiply-shift map of a 32-bit word onto [0, bound).
__host__ __device__ __forceinline__ uint32_t mulhi_bound(uint32_t w, uint32_t bound) {
    return (uint32_t)(((uint64_t)w * (uint64_t)bound) >> 32);
}

}  // namespace mp

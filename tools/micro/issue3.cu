// Issue / pipe probe, round 2: throughput of single instruction kinds and of pairs of kinds interleaved in one
// warp, on sm_100a.  Questions it answers (profiles/r2_issue3_probe.txt):
//   - is the packed fp32 pair (fma.rn.f32x2 -> FFMA2) issued at the rate of a scalar FFMA?  (screening kernel)
//   - what do the instructions of the v2 generator (IMAD, IMAD.WIDE, SHF, LOP3) cost beside a DFMA stream?
//   - shared-memory loads/stores of 4, 8 and 16 bytes per lane
// Every companion has a loop-carried dependence through its own registers and its result is folded into the
// value the kernel may store, so ptxas can neither merge nor drop it (the SASS is checked: tools/micro/README).
// clock64 counts SM cycles; numbers are cycles per loop repetition (8 main + K companion instructions per warp),
// for 1, 2 and 4 warps per sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue3 issue3.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Kind { NONE, DFMA, DADD, FFMA, FFMA2, FADD2, FMUL2, FADD, IMAD, IMADW, IMADHI, LOP3, SHF, IADD3, LDS32, LDS64, LDS128, STS32, STS64, MUFURCP, F2I, I2FD, HILO };

struct Regs {
    double d; float f; unsigned long long p;  // p: packed f32x2
    unsigned u, v; unsigned long long w; float4 q;
};

template <int KIND>
__device__ __forceinline__ void op(Regs& r, unsigned sm_addr, double y, double z) {
    if (KIND == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r.d) : "d"(y), "d"(z));
    if (KIND == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(r.d) : "d"(z));
    if (KIND == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r.f) : "f"((float)y), "f"((float)z));
    if (KIND == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r.f) : "f"((float)z));
    if (KIND == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r.p) : "l"(r.w), "l"((unsigned long long)__double_as_longlong(z)));
    if (KIND == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r.p) : "l"(r.w));
    if (KIND == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r.p) : "l"(r.w));
    if (KIND == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r.u) : "r"(r.v));
    if (KIND == IMADW) asm volatile("{ .reg .u32 lo, hi; mul.wide.u32 %0, %1, %2; mov.b64 {lo, hi}, %0; xor.b32 %1, lo, hi; }" : "=l"(r.w), "+r"(r.u) : "r"(r.v));
    if (KIND == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r.u) : "r"(r.v));
    if (KIND == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r.u) : "r"(r.v), "r"((unsigned)r.w));
    if (KIND == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r.u) : "r"(r.v));
    if (KIND == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r.u) : "r"(r.v));
    if (KIND == LDS32) asm volatile("{ .reg .u32 t; ld.shared.u32 t, [%1]; add.u32 %0, %0, t; }" : "+r"(r.u) : "r"(sm_addr));
    if (KIND == LDS64) asm volatile("{ .reg .u64 t; ld.shared.u64 t, [%1]; xor.b64 %0, %0, t; }" : "+l"(r.w) : "r"(sm_addr));
    if (KIND == LDS128) asm volatile("{ .reg .f32 a, b, c, e; ld.shared.v4.f32 {a, b, c, e}, [%1]; add.f32 %0, %0, a; }" : "+f"(r.q.x) : "r"(sm_addr));
    if (KIND == STS32) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sm_addr), "r"(r.u));
    if (KIND == STS64) asm volatile("st.shared.u64 [%0], %1;" :: "r"(sm_addr), "l"(r.w));
    if (KIND == MUFURCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r.f));
    if (KIND == F2I) asm volatile("{ .reg .f32 t; cvt.rn.f32.u32 t, %0; mov.b32 %0, t; }" : "+r"(r.u));
    if (KIND == I2FD) asm volatile("{ .reg .f64 t; cvt.rn.f64.u64 t, %1; add.rn.f64 %0, %0, t; }" : "+d"(r.d) : "l"(r.w));
    if (KIND == HILO) asm volatile("{ .reg .f64 t; mov.b64 t, {%1, %2}; add.rn.f64 %0, %0, t; }" : "+d"(r.d) : "r"(r.u), "r"(r.v));
}

template <int MAIN, int NM, int COMP, int NC>
__global__ void __launch_bounds__(512, 1) body(double* out, int iters, double y, double z, long long* cyc) {
    __shared__ __align__(16) double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1.0 + i;
    __syncthreads();
    const unsigned sm_addr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    Regs a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i].d = threadIdx.x * 1e-9 + i; a[i].f = 1.0f + i; a[i].p = 0x3f8000003f800000ull; a[i].u = threadIdx.x + i; a[i].v = 0x7fffffffu - 2 * i;
        a[i].w = 0x3f7ffff03f7ffff0ull + i + (threadIdx.x & 3); a[i].q = make_float4(0, 0, 0, 0);
        b[i] = a[i]; b[i].u += 77;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < NM) op<MAIN>(a[i], sm_addr, y, z);
                if (i < NC) op<COMP>(b[i], sm_addr + 8192, y, z);
            }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].d + a[i].f + (double)a[i].p + a[i].u + (double)a[i].w + a[i].q.x + b[i].d + b[i].f + (double)b[i].p + b[i].u + (double)b[i].w + b[i].q.x;
    if (s == 12345.6789) out[0] = s;
}

template <int MAIN, int NM, int COMP, int NC>
static void run(const char* name, double* out, long long* cyc) {
    const int iters = 1024;
    printf("%-34s", name);
    for (int wps : {1, 2, 4}) {
        for (int r = 0; r < 2; ++r) { body<MAIN, NM, COMP, NC><<<148, 128 * wps>>>(out, iters, 0.9999999, 1e-9, cyc); cudaDeviceSynchronize(); }
        printf("  w%d: %7.2f", wps, (double)cyc[0] / (iters * 4.0));
    }
    printf("\n");
}

#define SOLO(K) run<K, 8, NONE, 0>("8 " #K, out, cyc);
#define PAIR(M, C) run<M, 8, C, 4>("8 " #M " + 4 " #C, out, cyc); run<M, 8, C, 8>("8 " #M " + 8 " #C, out, cyc);

int main() {
    double* out; cudaMalloc(&out, 64);
    long long* cyc; cudaMallocManaged(&cyc, 64);
    printf("SM cycles per repetition (clock64); w = warps per sub-partition\n");
    SOLO(DFMA) SOLO(DADD) SOLO(FFMA) SOLO(FADD) SOLO(FFMA2) SOLO(FADD2) SOLO(FMUL2) SOLO(IMAD) SOLO(IMADW) SOLO(IMADHI) SOLO(LOP3) SOLO(SHF) SOLO(IADD3)
    SOLO(LDS32) SOLO(LDS64) SOLO(LDS128) SOLO(STS32) SOLO(STS64) SOLO(MUFURCP) SOLO(F2I) SOLO(I2FD) SOLO(HILO)
    PAIR(DFMA, IMAD) PAIR(DFMA, IMADW) PAIR(DFMA, LOP3) PAIR(DFMA, SHF) PAIR(DFMA, IADD3) PAIR(DFMA, LDS64) PAIR(DFMA, STS64) PAIR(DFMA, FFMA) PAIR(DFMA, I2FD) PAIR(DFMA, HILO)
    PAIR(FFMA2, LDS64) PAIR(FFMA2, LDS32) PAIR(FFMA2, IMAD) PAIR(FFMA2, LOP3) PAIR(FFMA2, SHF) PAIR(FFMA2, FFMA) PAIR(FFMA2, MUFURCP) PAIR(FFMA2, STS32) PAIR(FFMA2, IMADW)
    PAIR(FFMA, LDS32) PAIR(FFMA, IMAD) PAIR(FFMA, LOP3)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

// Philox4x32-10 latency/throughput probe and cross-warp interference with an FP64 stream.
// roles per warp (warp w runs on sub-partition w % 4): 'P' = dependent Philox chains (ILP chains per
// thread), 'Q' = the same with the products split into mul.hi + mul.lo, 'D' = 8 independent DFMA chains,
// 'M'/'N' = one Philox call + 40/80 DFMAs interleaved in one warp, '-' = exit at once.  Prints cycles per Philox call
// (per chain) and cycles per DFMA for every configuration.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../maxpro_b200/csrc -o philox_lat philox_lat.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "philox.cuh"
using namespace mp;

// Philox round with the 32x32->64 product taken as mul.hi + mul.lo instead of one mul.wide
__device__ __forceinline__ Philox4 philox_hilo(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t m0, uint32_t m1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi0) : "r"(c0));
        lo0 = c0 * m0;
        asm volatile("mul.hi.u32 %0, %1, 0xCD9E8D57;" : "=r"(hi1) : "r"(c2));
        lo1 = c2 * m1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c1 = lo1; c3 = lo0; c0 = n0; c2 = n2;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return Philox4{c0, c1, c2, c3};
}
template <int ILP>
__device__ __forceinline__ unsigned philox2_role(int iters, uint32_t m0, uint32_t m1) {
    uint32_t c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            Philox4 w = philox_hilo(c[i], 1u, 2u, 3u, kKeyLhd0, kKeyLhd1, m0, m1);
            c[i] = w.x ^ w.y ^ w.z ^ w.w;
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    return s;
}

template <int ILP>
__device__ __forceinline__ unsigned philox_role(int iters) {
    uint32_t c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            Philox4 w = philox4x32_10(c[i], 1u, 2u, 3u, kKeyLhd0, kKeyLhd1);
            c[i] = w.x ^ w.y ^ w.z ^ w.w;
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    return s;
}

__device__ __forceinline__ double dfma_role(int iters, double y, double z) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], y, z);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    return s;
}

template <int ILP, int NDF>
__device__ __forceinline__ double mixed_role(int iters, double y, double z) {
    uint32_t c[ILP];
    double x[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 7 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            Philox4 w = philox4x32_10(c[i], 1u, 2u, 3u, kKeyLhd0, kKeyLhd1);
            c[i] = w.x ^ w.y ^ w.z ^ w.w;
#pragma unroll
            for (int r = 0; r < NDF / 8; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fma(x[j], y, z);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    return s;
}

// roles: string of up to 16 chars, one per warp
template <int ILP>
__global__ void probe(const char* roles, int nw, int p_iters, int d_iters, double y, double z, long long* cyc, double* sink, uint32_t m0, uint32_t m1) {
    const int w = threadIdx.x >> 5;
    const char role = roles[w];
    __syncthreads();
    long long t0 = clock64();
    double s = 0;
    if (role == 'P') s = philox_role<ILP>(p_iters);
    else if (role == 'Q') s = philox2_role<ILP>(p_iters, m0, m1);
    else if (role == 'D') s = dfma_role(d_iters, y, z);
    else if (role == 'M') s = mixed_role<ILP, 40>(p_iters, y, z);
    else if (role == 'N') s = mixed_role<ILP, 80>(p_iters, y, z);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[w] = t1 - t0;
    if (s == 123.456) sink[0] = s;
}

template <int ILP>
static void run(const char* roles) {
    const int nw = (int)strlen(roles);
    char* d_roles; cudaMalloc(&d_roles, 32); cudaMemcpy(d_roles, roles, nw + 1, cudaMemcpyHostToDevice);
    long long* cyc; cudaMallocManaged(&cyc, 32 * 8);
    double* sink; cudaMalloc(&sink, 8);
    const int p_iters = 4096 / ILP, d_iters = 4096;
    for (int r = 0; r < 2; ++r) {
        probe<ILP><<<148, 32 * nw>>>(d_roles, nw, p_iters, d_iters, 0.9999999, 1e-9, cyc, sink, kPhiloxM0, kPhiloxM1);
        cudaDeviceSynchronize();
    }
    printf("roles %-16s ILP %d:", roles, ILP);
    for (int w = 0; w < nw; ++w) {
        if (roles[w] == 'P' || roles[w] == 'Q') printf("  w%d P %.0f cyc/call (%.1f/call/chain-slot)", w, (double)cyc[w] / p_iters, (double)cyc[w] / (p_iters * ILP));
        else if (roles[w] == 'M' || roles[w] == 'N') printf("  w%d %c %.1f cyc per (call + %d DFMA)", w, roles[w], (double)cyc[w] / (p_iters * ILP), roles[w] == 'M' ? 40 : 80);
        else if (roles[w] == 'D') printf("  w%d D %.2f cyc/DFMA", w, (double)cyc[w] / (d_iters * 32.0));
    }
    printf("\n");
    cudaFree(d_roles); cudaFree(cyc); cudaFree(sink);
}

int main() {
    printf("warp w runs on sub-partition w %% 4; 'P' Philox4x32-10 chains, 'D' DFMA stream\n");
    printf("-- Philox alone: latency (ILP 1) and throughput\n");
    run<1>("P"); run<2>("P"); run<4>("P"); run<8>("P"); run<2>("P---P"); run<2>("P---P---P");
    printf("-- DFMA alone\n");
    run<1>("D"); run<1>("D---D"); run<1>("D---D---D");
    printf("-- a Philox warp beside DFMA warps on the same sub-partition\n");
    run<2>("P---D"); run<2>("P---D---D"); run<2>("P---P---D"); run<4>("P---D---D");
    printf("-- one call + 40 (M) or 80 (N) DFMAs interleaved in the same warp\n");
    run<2>("M"); run<2>("M---M---M"); run<4>("M---M---M"); run<2>("N"); run<2>("N---N---N");
    printf("-- 32x32->64 products as mul.hi + mul.lo with a register multiplier (no IMAD.WIDE)\n");
    run<2>("Q"); run<2>("Q---Q---Q"); run<2>("Q---D"); run<2>("Q---D---D");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

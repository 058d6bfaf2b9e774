// K1 alone: materialise designs of the LHD stream (v1 or v2, philox.cuh) in device memory.
//
// Replaces lhd::generate_lhd (src/lhd.rs:97-132) for the cases where a design
// has to exist outside a search kernel: maxpro_generate_lhd, build_lhd with
// metric None (src/lib.rs:46-50) and regeneration of a search winner from its
// (seed, index).  One CTA per (design, column); the inside-out Fisher-Yates is a
// chain of n dependent steps, so a single lane walks it in shared memory and the
// whole CTA writes the column out.  Not a throughput path: the search kernels
// generate their candidates in place and never call this.
#include "common.cuh"
#include "philox.cuh"
#include "kernels.h"

namespace mp {

__device__ __forceinline__ void fy_column(int gen, double* col, size_t stride, unsigned long long n, uint32_t k,
                                          uint32_t s_lo, uint32_t s_hi, double c1, double c0) {
    LhdKeyCache kc{0xFFFFFFFFu, 0u, 0u};
    for (unsigned long long i = 0; i < n; i += 2) {
        const LhdBlock w = lhd_draw_block(gen, (uint32_t)(i >> 1), k, s_lo, s_hi, kc);
        {
            const double v = lhd_value(gen, (uint32_t)i, w.jit0, c1, c0);
            col[i * stride] = col[w.j0 * stride];
            col[w.j0 * stride] = v;
        }
        if (i + 1 < n) {
            const double v = lhd_value(gen, (uint32_t)i + 1u, w.jit1, c1, c0);
            col[(i + 1) * stride] = col[w.j1 * stride];
            col[w.j1 * stride] = v;
        }
    }
}

__global__ void __launch_bounds__(128) generate_kernel(unsigned long long n, unsigned long long d,
                                                       unsigned long long seed, unsigned long long lo,
                                                       unsigned long long count, double* out, double c1,
                                                       double c0, int use_smem, int gen) {
    extern __shared__ double col_s[];
    for (unsigned long long item = blockIdx.x; item < count * d; item += gridDim.x) {
        const unsigned long long c = item / d, k = item % d;
        const unsigned long long stream = seed + lo + c;  // src/lib.rs:70
        double* dst = out + c * n * d + k;                 // lhd[i][k], row-major
        if (use_smem) {
            if (threadIdx.x == 0)
                fy_column(gen, col_s, 1, n, (uint32_t)k, (uint32_t)stream, (uint32_t)(stream >> 32), c1, c0);
            __syncthreads();
            for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) dst[i * d] = col_s[i];
            __syncthreads();
        } else if (threadIdx.x == 0) {
            fy_column(gen, dst, d, n, (uint32_t)k, (uint32_t)stream, (uint32_t)(stream >> 32), c1, c0);
        }
    }
}

cudaError_t launch_generate(unsigned long long n, unsigned long long d, unsigned long long seed,
                            unsigned long long lo, unsigned long long count, double* d_out,
                            cudaStream_t stream) {
    if (n == 0 || d == 0 || count == 0) return cudaSuccess;
    const int gen = stream_version();
    const double c1 = ldexp(1.0 / (double)n, -32), c0 = lhd_c0(gen, c1);
    size_t smem = n * sizeof(double);
    int use_smem = smem <= 200 * 1024;
    if (!use_smem) smem = 0;
    cudaError_t e = cudaFuncSetAttribute(generate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return e;
    unsigned long long items = count * d;
    int grid = (int)(items < 148ull * 64 ? items : 148ull * 64);
    generate_kernel<<<grid, 128, smem, stream>>>(n, d, seed, lo, count, d_out, c1, c0, use_smem, gen);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

// FP64 FMA-pipe peak probe: the roofline denominator for the criterion kernels.
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so bench.py measures the
// non-tensor FP64 peak itself: 8 independent DFMA chains per thread, enough
// warps per SM to cover the pipe latency, no memory traffic in the loop.
#include "common.cuh"
#include "kernels.h"

namespace mp {

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* sink, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
    double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.6789) sink[0] = s;  // never true; keeps the chains alive
}

cudaError_t launch_dfma_probe(double* d_sink, int grid, int threads, int iters, cudaStream_t stream) {
    dfma_probe_kernel<<<grid, threads, 0, stream>>>(d_sink, iters, 0.999999, 1e-7);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mp

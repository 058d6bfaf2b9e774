// FP64 FMA-pipe peak probe: the roofline denominator for the criterion kernels.
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so bench.py measures the
// non-tensor FP64 peak itself: 8 independent three-register DFMA chains per thread,
// enough warps per SM to cover the pipe latency, no memory traffic in the loop.
#include "common.cuh"
#include "kernels.h"

// ---- FP64 pipe probes: the peak probe behind maxpro_fp64_peak_probe, and a tuning aid ---------
namespace mp {

// OP: 0 fma(x,a,b) with uniform a,b   1 fma(x,y,z) three registers   2 x+y   3 x*y
//     4 the MaxPro pair body on register data (3 sub, 2 mul, fma, chain push)
template <int ILP, int OP>
__global__ void dp_chain_kernel(double* sink, int iters, double a, double b, const double* vals) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = vals[i]; z[i] = vals[i + 8]; }
    double num[ILP], den[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { num[i] = 0.0; den[i] = 1.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            else if (OP == 1) x[i] = fma(x[i], y[i], z[i]);
            else if (OP == 2) x[i] = x[i] + y[i];
            else if (OP == 3) x[i] = x[i] * y[i];
            else if (OP == 5 || OP == 6 || OP == 7) {
                // one DFMA + NI independent integer ops: does the FP64 issue shadow the ALU/FMA pipes?
                x[i] = fma(x[i], y[i], z[i]);
                unsigned int u = __double2loint(num[i]);
                if (OP >= 5) u = u * 1664525u + 1013904223u;          // IMAD (fma pipe)
                if (OP >= 6) u = (u ^ (u >> 7)) + 0x9E3779B9u;        // LOP3/SHF/IADD (alu pipe)
                if (OP >= 7) u = u * 22695477u + 1u;
                num[i] = __hiloint2double(__double2hiint(num[i]), (int)u);
            }
            else {
                double q = (x[i] - y[i]) * (y[i] - z[i]) * (z[i] - x[i]);
                double p = fma(q, q, 1e-12);
                num[i] = fma(num[i], p, den[i]);
                den[i] *= p;
                x[i] = den[i] > 1e-100 ? x[i] : x[i] + 1.0;  // keeps den live without a reset
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + num[i] + den[i];
    if (s == 12345.6789) sink[0] = s;
}

cudaError_t launch_dfma_probe(double* d_sink, int grid, int threads, int iters, cudaStream_t stream) {
    // d_sink: 17 doubles -- [0] sink, [1..16] operand values (set by the caller).
    // Eight independent x = fma(x, y, z) chains per thread, all operands in registers: the form
    // that reaches 1 warp-instruction / 2 cycles / sub-partition (37.1 TFLOP/s on a B200 at
    // 1965 MHz; the constant-operand form fma(x, c1, c2) tops out ~8% lower).
    dp_chain_kernel<8, 1><<<grid, threads, 0, stream>>>(d_sink, iters, 0.0, 0.0, d_sink + 1);
    count_launch();
    return cudaGetLastError();
}

// FP32 FMA-pipe peak probe (roofline denominator of the fp32 screening kernel): eight independent three-register
// FFMA chains per thread, no memory traffic in the loop.
__global__ void sp_chain_kernel(float* sink, int iters, const float* vals) {
    float x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-6f + i; y[i] = vals[i]; z[i] = vals[i + 8]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], y[i], z[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678f) sink[0] = s;
}

cudaError_t launch_ffma_probe(float* d_sink, int grid, int threads, int iters, cudaStream_t stream) {
    // d_sink: 17 floats -- [0] sink, [1..16] operand values (set by the caller)
    sp_chain_kernel<<<grid, threads, 0, stream>>>(d_sink, iters, d_sink + 1);
    count_launch();
    return cudaGetLastError();
}

template <int ILP>
static void launch_ilp(int op, int grid, int threads, double* sink, int iters, const double* vals) {
    switch (op) {
        case 0: dp_chain_kernel<ILP, 0><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 1: dp_chain_kernel<ILP, 1><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 2: dp_chain_kernel<ILP, 2><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 3: dp_chain_kernel<ILP, 3><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 5: dp_chain_kernel<ILP, 5><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 6: dp_chain_kernel<ILP, 6><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        case 7: dp_chain_kernel<ILP, 7><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
        default: dp_chain_kernel<ILP, 4><<<grid, threads>>>(sink, iters, 0.999999, 1e-7, vals); break;
    }
}

}  // namespace mp

extern "C" __attribute__((visibility("default"))) int maxpro_debug_dp_probe(int grid, int threads, int ilp, int op,
                                                                              int iters, double* ms_out) {
    double* sink = nullptr;
    double* vals = nullptr;
    cudaMalloc(&sink, 8);
    cudaMalloc(&vals, 16 * 8);
    double h[16];
    for (int i = 0; i < 16; ++i) h[i] = 0.999999 + 1e-7 * i;
    cudaMemcpy(vals, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        if (ilp == 1) mp::launch_ilp<1>(op, grid, threads, sink, iters, vals);
        else if (ilp == 2) mp::launch_ilp<2>(op, grid, threads, sink, iters, vals);
        else if (ilp == 4) mp::launch_ilp<4>(op, grid, threads, sink, iters, vals);
        else mp::launch_ilp<8>(op, grid, threads, sink, iters, vals);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaFree(sink); cudaFree(vals); cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out = best;
    return e == cudaSuccess ? 0 : 2;
}

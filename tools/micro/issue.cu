// Issue-model probe: does an FP64 instruction block the sub-partition's issue port for both of
// its two pipe cycles, and for which companion pipes (fma: IMAD, alu: SHF, lsu: LDS)?
// Each loop iteration issues NF independent DFMAs (register operands) plus NI IMADs, NL SHFs and
// NS LDS.64 on independent chains; time per iteration per sub-partition is printed in cycles for
// 1..4 warps per sub-partition.  Check the loop body with `cuobjdump -sass` before trusting it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue issue.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int NL, int NS>
__global__ void __launch_bounds__(512, 1) body(double* out, int iters, double y, double z, unsigned a, unsigned b, unsigned m) {
    __shared__ double sm[512 * 2];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 512] = 1.0;
    __syncthreads();
    double x[NF > 0 ? NF : 1];
    unsigned u[NI > 0 ? NI : 1], v[NL > 0 ? NL : 1];
    double ld[NS > 0 ? NS : 1];
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < NI; ++i) u[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NL; ++i) v[i] = threadIdx.x * 3 + i;
#pragma unroll
    for (int i = 0; i < NS; ++i) ld[i] = 0.0;
    const volatile double* base = sm + (threadIdx.x & 511);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < NF) x[i] = fma(x[i], y, z);
                if (i < NI) u[i] = u[i] * a + b;
                if (i < NL) v[i] = __funnelshift_l(v[i], v[(i + 1) % (NL > 0 ? NL : 1)], 7);
                if (i < NS) ld[i] = base[(i & 1) * 512];
            }
            if (NF > 8) {
#pragma unroll
                for (int i = 8; i < NF; ++i) x[i] = fma(x[i], y, z);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NI; ++i) s += u[i];
#pragma unroll
    for (int i = 0; i < NL; ++i) s += v[i];
#pragma unroll
    for (int i = 0; i < NS; ++i) s += ld[i];
    if (s == 12345.6789) out[0] = s;
}

template <int NF, int NI, int NL, int NS>
static void run(const char* name, double* out) {
    const int iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-28s", name);
    for (int wps = 1; wps <= 4; ++wps) {
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0);
            body<NF, NI, NL, NS><<<148, 128 * wps>>>(out, iters, 0.9999999, 1e-9, 1664525u, 1013904223u, 0x7fffff7fu);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double cyc = best * 1e-3 * 1.965e9 / (iters * 4.0);  // cycles per unrolled repetition per sub-partition
        printf("  w%d: %6.2f", wps, cyc);
    }
    printf("   (ideal fp64-only: %d)\n", 2 * NF);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    double* out; cudaMalloc(&out, 64);
    printf("cycles per repetition per sub-partition at 1965 MHz, warps per sub-partition 1..4 (all warps run the same body)\n");
    run<8, 0, 0, 0>("8 DFMA", out);
    run<0, 8, 0, 0>("8 IMAD", out);
    run<0, 0, 8, 0>("8 SHF", out);
    run<0, 0, 0, 8>("8 LDS", out);
    run<8, 4, 0, 0>("8 DFMA + 4 IMAD", out);
    run<8, 8, 0, 0>("8 DFMA + 8 IMAD", out);
    run<8, 0, 4, 0>("8 DFMA + 4 SHF", out);
    run<8, 0, 8, 0>("8 DFMA + 8 SHF", out);
    run<8, 0, 0, 2>("8 DFMA + 2 LDS", out);
    run<8, 0, 0, 4>("8 DFMA + 4 LDS", out);
    run<8, 0, 0, 8>("8 DFMA + 8 LDS", out);
    run<8, 4, 4, 0>("8 DFMA + 4 IMAD + 4 SHF", out);
    run<8, 8, 8, 0>("8 DFMA + 8 IMAD + 8 SHF", out);
    run<8, 4, 4, 2>("8 DFMA + 4 IMAD 4 SHF 2 LDS", out);
    run<0, 8, 8, 0>("8 IMAD + 8 SHF", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

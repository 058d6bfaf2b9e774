// Latency probe for the annealer's narrow half: dependent chains on one warp, cycles per op.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double wsum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void probe(double* out, long long* cyc, double x0, double y, int iters) {
    double x = x0 + threadIdx.x * 1e-9;
    long long t0, t1;
    int k = 0;
    // pow chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) x = pow(x, y) + 1e12;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // exp(log(x)*y)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) x = exp(log(x) * y) + 1e12;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // exp(-a/b)
    double b = 0.37 + x * 1e-14;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) b = exp(-b / 1.7) + 0.1;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // IEEE divide
    t0 = clock64();
    for (int i = 0; i < iters; ++i) b = 1.0 / b + 0.3;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // warp sum (double)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) b = wsum(b) * 0.03;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // __syncthreads
    t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // dependent DFMA
    t0 = clock64();
    for (int i = 0; i < iters; ++i) b = fma(b, 0.999, 0.001);
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    // sqrt
    t0 = clock64();
    for (int i = 0; i < iters; ++i) b = sqrt(b) + 0.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[k] = (t1 - t0) / iters; k++;
    out[threadIdx.x] = x + b;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096 * 8); cudaMallocManaged(&cyc, 64 * 8);
    const char* names[] = {"pow(x,0.05)+add", "exp(log(x)*y)+add", "exp(-b/c)+add", "1/b+add", "warp sum f64 + mul", "__syncthreads", "dfma", "sqrt+add"};
    for (int threads : {32, 64, 288, 544}) {
        probe<<<1, threads>>>(out, cyc, 5e11, 0.05, 200);
        cudaDeviceSynchronize();
        printf("threads=%d:", threads);
        for (int k = 0; k < 8; ++k) printf("  %s=%lld", names[k], cyc[k]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

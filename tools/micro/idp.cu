// IDP.4A (dp4a) issue rate and latency on one SM, next to IMAD and LDS: cycles per warp-instruction per sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o idp idp.cu && ./idp
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned int* out, long long* cyc, int iters) {
    unsigned int a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    const unsigned int b = out[0] | 0x01020304u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent dp4a chains
            a0 = __dp4a(a0, b, a0); a1 = __dp4a(a1, b, a1); a2 = __dp4a(a2, b, a2); a3 = __dp4a(a3, b, a3);
            a4 = __dp4a(a4, b, a4); a5 = __dp4a(a5, b, a5); a6 = __dp4a(a6, b, a6); a7 = __dp4a(a7, b, a7);
        } else if (MODE == 1) {  // one dependent chain
            a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0);
            a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0); a0 = __dp4a(a0, b, a0);
        } else {  // 8 independent IMAD chains
            a0 = a0 * b + a0; a1 = a1 * b + a1; a2 = a2 * b + a2; a3 = a3 * b + a3;
            a4 = a4 * b + a4; a5 = a5 * b + a5; a6 = a6 * b + a6; a7 = a7 * b + a7;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x + 1] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    unsigned int* out; long long* cyc;
    cudaMalloc(&out, 4 * 2048); cudaMemset(out, 0, 4 * 2048); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16}) {
        long long c[3];
        k<0><<<1, 32 * warps>>>(out, cyc, iters); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
        k<1><<<1, 32 * warps>>>(out, cyc, iters); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
        k<2><<<1, 32 * warps>>>(out, cyc, iters); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)(warps + 3) / 4;  // warps per sub-partition (rounded up for 1)
        printf("%2d warps: dp4a x8 independent %.2f cycles/instr/warp (%.2f per sub-partition slot), dependent chain %.2f cycles/instr, IMAD x8 independent %.2f cycles/instr/warp\n",
               warps, (double)c[0] / (8.0 * iters), (double)c[0] / (8.0 * iters) / (warps >= 4 ? warps / 4.0 : 1.0), (double)c[1] / (8.0 * iters), (double)c[2] / (8.0 * iters));
        (void)per;
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

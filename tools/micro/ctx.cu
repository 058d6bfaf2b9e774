// time of CUDA context creation alone (what every process that touches the GPU pays before its first kernel)
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    auto t0 = std::chrono::steady_clock::now();
    cudaFree(0);
    auto t1 = std::chrono::steady_clock::now();
    printf("cudaFree(0): %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
    return 0;
}

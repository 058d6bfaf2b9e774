// Same-warp co-issue probe, clock64-timed (true SM cycles): 8 independent DFMAs per repetition plus
// K companion instructions of one kind.  Companion kinds are written in inline PTX so that ptxas
// cannot fold them.  Prints cycles per repetition for 1 and 3 warps per sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue2 issue2.cu
#include <cstdio>
#include <cuda_runtime.h>

enum Kind { NONE, IMADW, IMADHI, IMAD, LOP3I, VIADD_, VMNMX, I2F64, LDS64, STS64, RCP64H, SHFL_ };

template <int KIND>
__device__ __forceinline__ void companion(unsigned& u, unsigned& v, unsigned long long& w, double& f, unsigned sm_addr) {
    if (KIND == IMADW) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w) : "r"((unsigned)w));
    if (KIND == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(u) : "r"(v));
    if (KIND == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u) : "r"(v));
    if (KIND == LOP3I) asm volatile("lop3.b32 %0, %0, 0x9E3779B9, %1, 0x96;" : "+r"(u) : "r"(v));
    if (KIND == VIADD_) asm volatile("add.u32 %0, %0, 0x600;" : "+r"(u));
    if (KIND == VMNMX) asm volatile("{ .reg .u32 t; add.u32 t, %0, 0x600; min.u32 %0, t, %1; }" : "+r"(u) : "r"(v));
    if (KIND == I2F64) asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(f) : "l"(w));
    if (KIND == LDS64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f) : "r"(sm_addr));
    if (KIND == STS64) asm volatile("st.shared.f64 [%0], %1;" :: "r"(sm_addr), "d"(f));
    if (KIND == RCP64H) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(f));
    if (KIND == SHFL_) asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 0x1f, 0xffffffff;" : "+r"(u));
}

template <int KIND, int K, int NF>
__global__ void __launch_bounds__(512, 1) body(double* out, int iters, double y, double z, long long* cyc) {
    __shared__ double sm[1024];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 512] = 1.0;
    __syncthreads();
    const unsigned sm_addr = (unsigned)__cvta_generic_to_shared(sm + threadIdx.x);
    double x[8];
    unsigned u[8], v[8]; unsigned long long w[8]; double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-9 + i; u[i] = threadIdx.x + i; v[i] = 0x7fffffffu - i; w[i] = threadIdx.x * 77 + i; f[i] = 1.0 + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < NF) x[i] = fma(x[i], y, z);
                if (i < K) companion<KIND>(u[i], v[i], w[i], f[i], sm_addr + (i & 1) * 4096);
            }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + u[i] + v[i] + (double)w[i] + f[i];
    if (s == 12345.6789) out[0] = s;
}

template <int KIND, int K, int NF>
static void run(const char* name, double* out, long long* cyc) {
    const int iters = 2048;
    printf("%-30s", name);
    for (int wps : {1, 3}) {
        for (int r = 0; r < 2; ++r) { body<KIND, K, NF><<<148, 128 * wps>>>(out, iters, 0.9999999, 1e-9, cyc); cudaDeviceSynchronize(); }
        printf("  w%d: %7.2f", wps, (double)cyc[0] / (iters * 4.0));
    }
    printf("\n");
}

#define RUNK(KIND, label) \
    run<KIND, 8, 0>("0 DFMA + 8 " label, out, cyc); \
    run<KIND, 4, 8>("8 DFMA + 4 " label, out, cyc); \
    run<KIND, 8, 8>("8 DFMA + 8 " label, out, cyc);

int main() {
    double* out; cudaMalloc(&out, 64);
    long long* cyc; cudaMallocManaged(&cyc, 64);
    printf("true SM cycles per repetition (clock64), warps per sub-partition 1 and 3\n");
    run<NONE, 0, 8>("8 DFMA", out, cyc);
    RUNK(IMADW, "IMAD.WIDE") RUNK(IMADHI, "IMAD.HI") RUNK(IMAD, "IMAD") RUNK(LOP3I, "LOP3 imm") RUNK(VIADD_, "VIADD")
    RUNK(VMNMX, "VIADD+VIMNMX") RUNK(I2F64, "I2F.F64.U64") RUNK(LDS64, "LDS.64") RUNK(STS64, "STS.64") RUNK(RCP64H, "MUFU.RCP64H") RUNK(SHFL_, "SHFL")
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

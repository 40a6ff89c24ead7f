// Microbenchmark: issue cost of packed FFMA2 (fma.rn.f32x2) against scalar FFMA, alone and next to MUFU.EX2, per SM
// sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NF2, int NM>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[8], y[8];
    unsigned long long z[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); y[i] = 0.5f + i; z[i] = 0x3f0000003f000000ull + i; }
    const float a = 0.999f, b = 1e-3f;
    const unsigned long long a2 = 0x3f7fbe773f7fbe77ull, b2 = 0x3a83126f3a83126full;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
            for (int j = 0; j < NF; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[(i + j) & 7]) : "f"(a), "f"(b));
#pragma unroll
            for (int j = 0; j < NF2; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(z[(i + j) & 7]) : "l"(a2), "l"(b2));
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (float)(z[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NF2, int NM>
void run(int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<NF, NF2, NM><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<NF, NF2, NM><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  per iter: MUFU %d  FFMA %2d  FFMA2 %2d : %7.1f clk\n", warps, NM, 8 * NF, 8 * NF2, (double)c / iters);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8}) {
        run<8, 0, 0>(w); run<0, 8, 0>(w); run<0, 4, 0>(w); run<4, 4, 0>(w); run<0, 4, 8>(w); run<4, 0, 8>(w); run<0, 2, 8>(w); run<2, 0, 8>(w);
    }
    return 0;
}

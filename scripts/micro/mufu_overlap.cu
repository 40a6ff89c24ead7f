// Microbenchmark: can MUFU.EX2 (XU pipe) overlap with FFMA issue on one SM sub-partition?  Per loop iteration and thread:
// 8 independent ex2 and 8*K independent FFMAs.  Prints cycles per iteration for 1 and 2 warps per quadrant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_overlap mufu_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int M>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[8], y[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); y[i] = 0.5f + i; }
    const float a = 0.999f, b = 1e-3f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < M) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
            for (int j = 0; j < K; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[(i + j) & 7]) : "f"(a), "f"(b));
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int K, int M>
void run(int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<K, M><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<K, M><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps/SM %2d  MUFU %d  FFMA %2d per iter:  %7.1f clk/iter\n", warps, M, 8 * K, (double)c / iters);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8}) {
        run<0, 8>(w); run<1, 8>(w); run<2, 8>(w); run<4, 8>(w); run<8, 8>(w); run<4, 0>(w); run<8, 0>(w);
    }
    return 0;
}

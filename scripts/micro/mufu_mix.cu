// Microbenchmark: which instruction classes overlap with MUFU.EX2 on one SM sub-partition (B200)?  Per iteration and thread:
// 8 MUFU (or none) + 32 instructions of one class.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_mix mufu_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
enum { FFMA, FFMA2, FADD2, FMNMX, F2FP, IADD, HFMA2, FMNMX3, LOP, IMAD, SHFL };
template <int OP> __device__ __forceinline__ void op(float& y, unsigned long long& z, unsigned& u, float a, float b, unsigned long long a2, unsigned long long b2) {
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(a), "f"(b));
    if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(z) : "l"(a2), "l"(b2));
    if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(z) : "l"(b2));
    if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(y) : "f"(a));
    if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(a), "f"(b));
    if (OP == F2FP) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(y), "f"(a));
    if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(u) : "r"(0x1234567));
    if (OP == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u) : "r"(0x1234567));
    if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u) : "r"(0x800001), "r"(77));
    if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(0x3c003c00), "r"(0x00010001));
    if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(u));
}
template <int OP, int NM>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[8], y[8]; unsigned long long z[8]; unsigned u[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); y[i] = 0.5f + i; z[i] = 0x3f0000003f000000ull + i; u[i] = threadIdx.x + i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
            for (int j = 0; j < 4; ++j) op<OP>(y[(i + j) & 7], z[(i + j) & 7], u[(i + j) & 7], 0.999f, 1e-3f, 0x3f7fbe773f7fbe77ull, 0x3a83126f3a83126full);
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + y[i] + (float)(z[i] & 0xff) + (float)(u[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000; long long c0, c1;
    for (int w : {4, 8}) {
        k<OP, 0><<<148, w * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k<OP, 0><<<148, w * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        k<OP, 8><<<148, w * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k<OP, 8><<<148, w * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-7s warps/SM %d: 32 ops alone %6.1f clk | with 8 MUFU (%d clk alone) %6.1f clk\n", name, w, (double)c0 / iters, w == 4 ? 65 : 128, (double)c1 / iters);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<FFMA>("FFMA"); run<FFMA2>("FFMA2"); run<FADD2>("FADD2"); run<FMNMX>("FMNMX"); run<FMNMX3>("FMNMX3"); run<F2FP>("F2FP");
    run<IADD>("IADD"); run<LOP>("LOP"); run<IMAD>("IMAD"); run<HFMA2>("HFMA2"); run<SHFL>("SHFL");
    return 0;
}

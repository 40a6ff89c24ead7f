// 3x3 convolution with Cin <= 4 (NHWC fp16, 4 channels) on CUDA cores: the conv_in layers of the UNet, ControlNet,
// ControlNet conditioning embedding and VAE (K = 36 is far too small for a tensor-core tile to pay off and the op is
// bound by the output write).  One thread = one output pixel; the 3x3x4 patch lives in registers, weights in shared
// memory (broadcast reads), 8 output channels per 128-bit store.
#include "fie_common.cuh"

namespace fie {

__global__ void __launch_bounds__(128) k_conv_cin4(const uint2* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                                                   __half* __restrict__ out, int ld_out, int n, int h, int w, int cout, int act) {
    extern __shared__ float sw[];   // [cout][36] + bias[cout]
    for (int i = threadIdx.x; i < cout * 36; i += blockDim.x) sw[i] = wgt[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cout * 36 + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const long long npix = (long long)n * h * w;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const int xx = (int)(pix % w); const int yy = (int)((pix / w) % h); const long long img = pix / ((long long)w * h);
    float patch[36];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int y = yy + t / 3 - 1, xq = xx + t % 3 - 1;
        if (y >= 0 && y < h && xq >= 0 && xq < w) {
            uint2 u = __ldg(x + (img * h + y) * w + xq);
            float2 a = __half22float2(*reinterpret_cast<__half2*>(&u.x)), b = __half22float2(*reinterpret_cast<__half2*>(&u.y));
            patch[4 * t] = a.x; patch[4 * t + 1] = a.y; patch[4 * t + 2] = b.x; patch[4 * t + 3] = b.y;
        } else { patch[4 * t] = patch[4 * t + 1] = patch[4 * t + 2] = patch[4 * t + 3] = 0.f; }
    }
    __half* o = out + pix * ld_out;
    if (ld_out & 7) {   // narrow outputs (e.g. 4 channels): 64-bit stores
        for (int co = 0; co < ld_out; co += 4) {
            float acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = 0.f;
                if (co + j < cout) {
                    const float* wr = sw + (co + j) * 36;
                    s = sw[cout * 36 + co + j];
#pragma unroll
                    for (int k = 0; k < 36; ++k) s = fmaf(patch[k], wr[k], s);
                    if (act == FIE_ACT_SILU) s = silu_f(s);
                }
                acc[j] = s;
            }
            uint2 u; __half2* hh = reinterpret_cast<__half2*>(&u);
            hh[0] = __floats2half2_rn(acc[0], acc[1]); hh[1] = __floats2half2_rn(acc[2], acc[3]);
            *reinterpret_cast<uint2*>(o + co) = u;
        }
        return;
    }
    for (int co = 0; co < ld_out; co += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = 0.f;
            if (co + j < cout) {
                const float* wr = sw + (co + j) * 36;
                s = sw[cout * 36 + co + j];
#pragma unroll
                for (int k = 0; k < 36; ++k) s = fmaf(patch[k], wr[k], s);
                if (act == FIE_ACT_SILU) s = silu_f(s);
            }
            acc[j] = s;
        }
        uint4 u; __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) hh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        *reinterpret_cast<uint4*>(o + co) = u;
    }
}

}  // namespace fie
using namespace fie;

extern "C" int fie_conv3x3_cin4_f16(const void* x, const float* wgt, const float* bias, void* out, int ld_out,
                                    int n, int h, int w, int cout, int act, void* stream) {
    FIE_REQUIRE(x && wgt && out && n > 0 && h > 0 && w > 0 && cout > 0, "fie_conv3x3_cin4_f16: bad args");
    FIE_REQUIRE(ld_out >= cout && (ld_out % 4) == 0, "fie_conv3x3_cin4_f16: ld_out must be >= cout and a multiple of 4");
    const size_t smem = (size_t)cout * 37 * sizeof(float);
    FIE_REQUIRE(smem <= 96 * 1024, "fie_conv3x3_cin4_f16: cout too large");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_conv_cin4, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    const long long npix = (long long)n * h * w;
    k_conv_cin4<<<(unsigned)((npix + 127) / 128), 128, smem, (cudaStream_t)stream>>>((const uint2*)x, wgt, bias, (__half*)out, ld_out, n, h, w, cout, act);
    return check_launch("fie_conv3x3_cin4_f16");
}

// 3x3 convolution with Cin <= 4 (NHWC fp16, 4 channels) on CUDA cores: the conv_in layers of the UNet, ControlNet,
// ControlNet conditioning embedding and VAE (K = 36 is far too small for a tensor-core tile to pay off; the op is
// bound by the output write).  A thread computes 8 consecutive output channels of 4 consecutive pixels (128-bit stores);
// the threads of a pixel are adjacent lanes, so a warp writes whole contiguous output rows.  Weights live in shared
// memory as [36][cout] so that a warp reads consecutive words (no bank conflicts, broadcast across pixels).
#include "fie_common.cuh"

namespace fie {

constexpr int CIN4_PX = 4;   // consecutive output pixels (along x) per thread: every weight vector read from smem feeds 4 pixels

__global__ void __launch_bounds__(256) k_conv_cin4(const uint2* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                                                   __half* __restrict__ out, int ld_out, int n, int h, int w, int cout, int act, int groups, int vec) {
    extern __shared__ float sw[];   // [36][cout_pad] + bias[cout_pad], cout_pad = groups * vec
    const int cpad = groups * vec;
    for (int i = threadIdx.x; i < 36 * cpad; i += blockDim.x) { const int k = i / cpad, co = i % cpad; sw[i] = co < cout ? wgt[co * 36 + k] : 0.f; }
    for (int i = threadIdx.x; i < cpad; i += blockDim.x) sw[36 * cpad + i] = (bias && i < cout) ? bias[i] : 0.f;
    __syncthreads();
    const int wq = (w + CIN4_PX - 1) / CIN4_PX;                // pixel quads per row
    const long long nquads = (long long)n * h * wq;
    const int qpb = blockDim.x / groups;                       // quads per block
    const int g = threadIdx.x % groups, ql = threadIdx.x / groups;
    if (ql >= qpb) return;
    for (long long quad = (long long)blockIdx.x * qpb + ql; quad < nquads; quad += (long long)gridDim.x * qpb) {
        const int xq0 = (int)(quad % wq) * CIN4_PX; const long long rowi = quad / wq; const int yy = (int)(rowi % h); const long long img = rowi / h;
        float in[3][CIN4_PX + 2][4];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int y = yy + ky - 1;
#pragma unroll
            for (int cx = 0; cx < CIN4_PX + 2; ++cx) {
                const int xx = xq0 + cx - 1;
                float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
                if (y >= 0 && y < h && xx >= 0 && xx < w) {
                    const uint2 u = __ldg(x + (img * h + y) * w + xx);
                    a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)); b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                }
                in[ky][cx][0] = a.x; in[ky][cx][1] = a.y; in[ky][cx][2] = b.x; in[ky][cx][3] = b.y;
            }
        }
        float acc[CIN4_PX][8];
#pragma unroll
        for (int px = 0; px < CIN4_PX; ++px)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[px][j] = (j < vec) ? sw[36 * cpad + g * vec + j] : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4* wr = reinterpret_cast<const float4*>(sw + (4 * t + c) * cpad + g * vec);
                const float4 w0 = wr[0];
                float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (vec == 8) w1 = wr[1];
#pragma unroll
                for (int px = 0; px < CIN4_PX; ++px) {
                    const float v = in[t / 3][px + t % 3][c];
                    acc[px][0] = fmaf(v, w0.x, acc[px][0]); acc[px][1] = fmaf(v, w0.y, acc[px][1]);
                    acc[px][2] = fmaf(v, w0.z, acc[px][2]); acc[px][3] = fmaf(v, w0.w, acc[px][3]);
                    if (vec == 8) {
                        acc[px][4] = fmaf(v, w1.x, acc[px][4]); acc[px][5] = fmaf(v, w1.y, acc[px][5]);
                        acc[px][6] = fmaf(v, w1.z, acc[px][6]); acc[px][7] = fmaf(v, w1.w, acc[px][7]);
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < CIN4_PX; ++px) {
            if (xq0 + px >= w) break;
            if (act == FIE_ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < vec && g * vec + j < cout) acc[px][j] = silu_f(acc[px][j]);
            }
            __half* o = out + ((img * h + yy) * w + xq0 + px) * ld_out + g * vec;
            if (vec == 8) {
                uint4 u; __half2* hh = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) hh[j] = __floats2half2_rn(acc[px][2 * j], acc[px][2 * j + 1]);
                *reinterpret_cast<uint4*>(o) = u;
            } else {
                uint2 u; __half2* hh = reinterpret_cast<__half2*>(&u);
                hh[0] = __floats2half2_rn(acc[px][0], acc[px][1]); hh[1] = __floats2half2_rn(acc[px][2], acc[px][3]);
                *reinterpret_cast<uint2*>(o) = u;
            }
        }
    }
}

}  // namespace fie
using namespace fie;

extern "C" int fie_conv3x3_cin4_f16(const void* x, const float* wgt, const float* bias, void* out, int ld_out,
                                    int n, int h, int w, int cout, int act, void* stream) {
    FIE_REQUIRE(x && wgt && out && n > 0 && h > 0 && w > 0 && cout > 0, "fie_conv3x3_cin4_f16: bad args");
    FIE_REQUIRE(ld_out >= cout && (ld_out % 4) == 0, "fie_conv3x3_cin4_f16: ld_out must be >= cout and a multiple of 4");
    const int vec = (ld_out % 8) == 0 ? 8 : 4;
    const int groups = ld_out / vec;                 // threads per pixel; channels >= cout are written as zeros
    FIE_REQUIRE(groups <= 256, "fie_conv3x3_cin4_f16: ld_out too large");
    const size_t smem = (size_t)37 * groups * vec * sizeof(float);
    FIE_REQUIRE(smem <= 160 * 1024, "fie_conv3x3_cin4_f16: cout too large");
    static bool attr_dev[kMaxDevices] = {false};
    bool& attr = attr_dev[current_device()];
    if (!attr) { cudaFuncSetAttribute(k_conv_cin4, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
    const long long nquads = (long long)n * h * ((w + CIN4_PX - 1) / CIN4_PX);
    const int qpb = 256 / groups;
    long long blocks = (nquads + qpb - 1) / qpb;
    const long long cap = 148ll * 8; if (blocks > cap) blocks = cap;
    k_conv_cin4<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>((const uint2*)x, wgt, bias, (__half*)out, ld_out, n, h, w, cout, act, groups, vec);
    return check_launch("fie_conv3x3_cin4_f16");
}

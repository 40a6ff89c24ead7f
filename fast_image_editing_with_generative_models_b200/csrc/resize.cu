// Lanczos-3 resampling of 8-bit RGB images, bit-identical to Pillow's Image.resize(size, Image.LANCZOS) (the host-side step at
// reference src/pipeline.py:251): two separable integer passes.  The per-output-coordinate windows (first input index, count) and
// 2^22 fixed-point weights are computed once on the host in double precision exactly as Pillow's precompute_coeffs /
// normalize_coeffs_8bpc do (fast_image_editing_with_generative_models_b200/resize.py); each pass accumulates in int32 from 2^21
// and clips (acc >> 22) to uint8.  One thread per output pixel (3 channels); the tables are tiny and stay in L1/L2.
#include "fie_common.cuh"

namespace fie {

constexpr int RS_PRECISION_BITS = 22;

__device__ __forceinline__ uint8_t rs_clip8(int v) { v >>= RS_PRECISION_BITS; return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// out[n, y, xx, c] = clip8(2^21 + sum_x in[n, y, x0 + x, c] * k[xx, x])
__global__ void __launch_bounds__(256) k_resample_h(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h, int w, int ow,
                                                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const long long total = (long long)n * h * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % ow); const long long row = i / ow;
        const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const int* k = kk + (long long)xx * ksize;
        const uint8_t* src = in + (row * w + x0) * 3;
        int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int x = 0; x < cnt; ++x) { const int c = __ldg(k + x); a0 += src[3 * x] * c; a1 += src[3 * x + 1] * c; a2 += src[3 * x + 2] * c; }
        uint8_t* d = out + i * 3;
        d[0] = rs_clip8(a0); d[1] = rs_clip8(a1); d[2] = rs_clip8(a2);
    }
}

// out[n, yy, x, c] = clip8(2^21 + sum_y in[n, y0 + y, x, c] * k[yy, y])
__global__ void __launch_bounds__(256) k_resample_v(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h, int w, int oh,
                                                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const long long total = (long long)n * oh * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w); const long long t = i / w; const int yy = (int)(t % oh); const long long img = t / oh;
        const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
        const int* k = kk + (long long)yy * ksize;
        const uint8_t* src = in + ((img * h + y0) * w + x) * 3;
        int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int y = 0; y < cnt; ++y) { const int c = __ldg(k + y); const uint8_t* s = src + (long long)y * w * 3; a0 += s[0] * c; a1 += s[1] * c; a2 += s[2] * c; }
        uint8_t* d = out + i * 3;
        d[0] = rs_clip8(a0); d[1] = rs_clip8(a1); d[2] = rs_clip8(a2);
    }
}

}  // namespace fie
using namespace fie;

extern "C" int fie_resample_lanczos_u8(const void* in, void* out, void* tmp, int n, int h, int w, int oh, int ow,
                                       const int* bounds_x, const int* coeff_x, int ksize_x, const int* bounds_y, const int* coeff_y, int ksize_y,
                                       void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "fie_resample_lanczos_u8: bad shape");
    const bool need_h = ow != w, need_v = oh != h;
    FIE_REQUIRE(!need_h || (bounds_x && coeff_x && ksize_x > 0), "fie_resample_lanczos_u8: horizontal tables missing");
    FIE_REQUIRE(!need_v || (bounds_y && coeff_y && ksize_y > 0), "fie_resample_lanczos_u8: vertical tables missing");
    FIE_REQUIRE(!(need_h && need_v) || tmp, "fie_resample_lanczos_u8: two passes need the [n,h,ow,3] temporary");
    auto grid = [](long long items) { long long b = (items + 255) / 256; const long long cap = 148ll * 16; return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b)); };
    if (!need_h && !need_v) {
        cudaError_t e = cudaMemcpyAsync(out, in, (size_t)n * h * w * 3, cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) { set_error("fie_resample_lanczos_u8: copy: %s", cudaGetErrorString(e)); return FIE_ERR_CUDA; }
        return FIE_OK;
    }
    const uint8_t* src = (const uint8_t*)in;
    if (need_h) {
        uint8_t* dst = need_v ? (uint8_t*)tmp : (uint8_t*)out;
        k_resample_h<<<grid((long long)n * h * ow), 256, 0, stream>>>(src, dst, n, h, w, ow, bounds_x, coeff_x, ksize_x);
        src = dst;
    }
    if (need_v) k_resample_v<<<grid((long long)n * oh * ow), 256, 0, stream>>>(src, (uint8_t*)out, n, h, ow, oh, bounds_y, coeff_y, ksize_y);
    return check_launch("fie_resample_lanczos_u8");
}

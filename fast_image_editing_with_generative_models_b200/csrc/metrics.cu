// Evaluation-metric kernels (reference src/metrics.py:150-381): everything around the three networks the
// reference's MetricsCalculator runs that is not a GEMM / LayerNorm / attention call of this library.
//   * SSIM (torchmetrics StructuralSimilarityIndexMeasure defaults: Gaussian 11x11, sigma 1.5, K = (0.01, 0.03)), exact
//     integer sum of squared differences (MSE / PSNR), on uint8 images
//   * the two float resamplers of the metric pre-processing: torchvision Resize(antialias=True) (DINO) in fp32
//     (the Pillow bicubic of the CLIP processor is integer and reuses fie_resample_lanczos_u8 with bicubic tables)
//   * ViT token assembly (patch rows, class token, position embeddings), row L2 normalisation, mean squared difference of two
//     fp32 matrices (DINO key self-similarity), cosine of row pairs (CLIPScore)
//   * LPIPS (SqueezeNet 1.1): explicit im2col for its odd-sized convolutions, ceil-mode 3x3/2 max pooling, and the per-layer
//     unit-normalise / squared-difference / 1x1 "lin" / spatial-mean reduction.
// All of it is HBM-bound byte / element work on images of at most a few MB; kernels are written for coalesced access and a grid
// that covers the 148 SMs, nothing more.
#include "fie_common.cuh"

namespace fie {

static inline unsigned mgrid(long long items, int block) {
    long long b = (items + block - 1) / block;
    const long long cap = 148ll * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ double block_sum_double(double v, double* sh) {      // result valid in thread 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)((blockDim.x + 31) >> 5) ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------------------------------
// SSIM: one CTA per 32 x 16 tile of the valid (h - 2R) x (w - 2R) map of one (image, channel); separable Gaussian of
// x, y, x^2, y^2, xy through shared memory, the SSIM map value per thread, per-image sums.  The pixels are the reference's float32
// v / 255; the moments are accumulated in DOUBLE: E[x^2] - E[x]^2 in float32 (what torchmetrics does) has a cancellation noise of ~1e-8
// against c2 = 9e-4, which moves the SSIM of flat regions by 1e-4 -- a few hundred DFMA per pixel cost nothing on 1.5 MB of input.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int SS_TW = 32, SS_TH = 16, SS_MAXK = 11;
struct SsimWeights { double w[SS_MAXK]; };

__global__ void __launch_bounds__(256) k_ssim_u8(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int h, int w, int c, int ks,
                                                 SsimWeights wt, double c1, double c2, double* __restrict__ out) {
    __shared__ float sa[SS_TH + SS_MAXK - 1][SS_TW + SS_MAXK - 1], sb[SS_TH + SS_MAXK - 1][SS_TW + SS_MAXK - 1];
    __shared__ double sh[5][SS_TH + SS_MAXK - 1][SS_TW];
    __shared__ double red[8];
    const int img = blockIdx.z / c, ch = blockIdx.z - img * c;
    const int ox0 = blockIdx.x * SS_TW, oy0 = blockIdx.y * SS_TH;
    const int th = SS_TH + ks - 1, tw = SS_TW + ks - 1, vh = h - ks + 1, vw = w - ks + 1;
    for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
        const int ty = i / tw, tx = i - ty * tw, y = oy0 + ty, x = ox0 + tx;
        float pa = 0.f, pb = 0.f;
        if (y < h && x < w) {
            const long long o = (((long long)img * h + y) * w + x) * c + ch;
            pa = (float)a[o] / 255.0f; pb = (float)b[o] / 255.0f;              // the reference's float32 pixels
        }
        sa[ty][tx] = pa; sb[ty][tx] = pb;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < th * SS_TW; i += blockDim.x) {
        const int ty = i / SS_TW, tx = i - ty * SS_TW;
        double s0 = 0., s1 = 0., s2 = 0., s3 = 0., s4 = 0.;
        for (int k = 0; k < ks; ++k) {
            const double g = wt.w[k], pa = (double)sa[ty][tx + k], pb = (double)sb[ty][tx + k];
            s0 = fma(g, pa, s0); s1 = fma(g, pb, s1); s2 = fma(g, pa * pa, s2); s3 = fma(g, pb * pb, s3); s4 = fma(g, pa * pb, s4);
        }
        sh[0][ty][tx] = s0; sh[1][ty][tx] = s1; sh[2][ty][tx] = s2; sh[3][ty][tx] = s3; sh[4][ty][tx] = s4;
    }
    __syncthreads();
    double local = 0.0;
    for (int i = threadIdx.x; i < SS_TH * SS_TW; i += blockDim.x) {
        const int ty = i / SS_TW, tx = i - ty * SS_TW;
        if (oy0 + ty >= vh || ox0 + tx >= vw) continue;
        double m[5] = {0., 0., 0., 0., 0.};
        for (int k = 0; k < ks; ++k) {
            const double g = wt.w[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fma(g, sh[q][ty + k][tx], m[q]);
        }
        const double mu_a2 = m[0] * m[0], mu_b2 = m[1] * m[1], mu_ab = m[0] * m[1];
        const double var_a = fmax(m[2] - mu_a2, 0.), var_b = fmax(m[3] - mu_b2, 0.), cov = m[4] - mu_ab;
        const double upper = 2. * cov + c2, lower = var_a + var_b + c2;
        local += ((2. * mu_ab + c1) * upper) / ((mu_a2 + mu_b2 + c1) * lower);
    }
    const double s = block_sum_double(local, red);
    if (threadIdx.x == 0) atomicAdd(out + img, s);
}

// sum over one image of (a - b)^2 on bytes: exact in 64-bit integers.  16 bytes per thread and load when the planes allow it:
// |a - b| per byte (vabsdiffu4) and its square-sum by dp4a, 32-bit partial sums per thread (<= 16 * 255^2 per iteration).
__global__ void __launch_bounds__(256) k_sqdiff_u8(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long per_image, int vec16,
                                                   unsigned long long* __restrict__ out) {
    const int img = blockIdx.y;
    const uint8_t* pa = a + (long long)img * per_image; const uint8_t* pb = b + (long long)img * per_image;
    unsigned long long acc = 0;
    if (vec16) {
        const uint4* va = reinterpret_cast<const uint4*>(pa); const uint4* vb = reinterpret_cast<const uint4*>(pb);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image / 16; i += (long long)gridDim.x * blockDim.x) {
            const uint4 x = __ldg(va + i), y = __ldg(vb + i);
            unsigned s = 0;
            unsigned d = __vabsdiffu4(x.x, y.x); s = __dp4a(d, d, s);
            d = __vabsdiffu4(x.y, y.y); s = __dp4a(d, d, s);
            d = __vabsdiffu4(x.z, y.z); s = __dp4a(d, d, s);
            d = __vabsdiffu4(x.w, y.w); s = __dp4a(d, d, s);
            acc += s;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (long long)gridDim.x * blockDim.x) {
            const int d = (int)pa[i] - (int)pb[i];
            acc += (unsigned)(d * d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + img, acc);
}

// ------------------------------------------------------------------------------------------------------------------------
// Float separable resampling with per-output windows (first index, count) and normalised fp32 weights: the antialiased bilinear
// Resize of torchvision / ATen (_upsample_bilinear2d_aa).  Horizontal pass reads uint8 (scaled by 1/255, as
// pil_to_tensor(img).float() / 255) or fp32, vertical pass applies (v - mean[c]) / std[c].
// ------------------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float rs_load(const T* p);
template <> __device__ __forceinline__ float rs_load<uint8_t>(const uint8_t* p) { return (float)(*p) / 255.0f; }
template <> __device__ __forceinline__ float rs_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float rs_load<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) k_resample_f32_h(const T* __restrict__ in, float* __restrict__ out, int n, int h, int w, int ow,
                                                        const int* __restrict__ bounds, const float* __restrict__ kk, int ksize) {
    const long long total = (long long)n * h * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % ow); const long long row = i / ow;
        const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
        const float* k = kk + (long long)xx * ksize;
        const T* src = in + (row * w + x0) * 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int x = 0; x < cnt; ++x) { const float g = __ldg(k + x); a0 = fmaf(rs_load(src + 3 * x), g, a0); a1 = fmaf(rs_load(src + 3 * x + 1), g, a1); a2 = fmaf(rs_load(src + 3 * x + 2), g, a2); }
        float* d = out + i * 3;
        d[0] = a0; d[1] = a1; d[2] = a2;
    }
}

struct Norm3 { float mean[3], inv_std[3]; };

template <typename T>
__global__ void __launch_bounds__(256) k_resample_f32_v(const T* __restrict__ in, float* __restrict__ out, int n, int h, int w, int oh,
                                                        const int* __restrict__ bounds, const float* __restrict__ kk, int ksize, Norm3 nm) {
    const long long total = (long long)n * oh * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w); const long long t = i / w; const int yy = (int)(t % oh); const long long img = t / oh;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        if (bounds) {
            const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
            const float* k = kk + (long long)yy * ksize;
            const T* src = in + ((img * h + y0) * w + x) * 3;
            for (int y = 0; y < cnt; ++y) { const float g = __ldg(k + y); const T* s = src + (long long)y * w * 3; a0 = fmaf(rs_load(s), g, a0); a1 = fmaf(rs_load(s + 1), g, a1); a2 = fmaf(rs_load(s + 2), g, a2); }
        } else {                                    // no vertical resampling (oh == h): normalisation only
            const T* s = in + ((img * h + yy) * w + x) * 3;
            a0 = rs_load(s); a1 = rs_load(s + 1); a2 = rs_load(s + 2);
        }
        float* d = out + i * 3;
        d[0] = (a0 - nm.mean[0]) * nm.inv_std[0]; d[1] = (a1 - nm.mean[1]) * nm.inv_std[1]; d[2] = (a2 - nm.mean[2]) * nm.inv_std[2];
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// ViT plumbing
// ------------------------------------------------------------------------------------------------------------------------
// Patch rows of a stride-P, kernel-P convolution as a GEMM operand: out[(img, gy, gx), (py, px, c)] = normalise(in[img, gy P + py, gx P + px, c]).
// One thread per output element pair would be byte-granular on the input; instead one thread per (row, py, px) pixel: 3 loads, 3 stores.
template <typename T>
__global__ void __launch_bounds__(256) k_patchify(const T* __restrict__ in, __half* __restrict__ out, int n, int h, int w, int P, Norm3 nm, int normalise) {
    const int gh = h / P, gw = w / P;
    const long long total = (long long)n * gh * gw * P * P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % P); long long t = i / P; const int py = (int)(t % P); t /= P;
        const int gx = (int)(t % gw); t /= gw; const int gy = (int)(t % gh); const long long img = t / gh;
        const T* s = in + ((img * h + (long long)gy * P + py) * w + (long long)gx * P + px) * 3;
        float v0 = rs_load(s), v1 = rs_load(s + 1), v2 = rs_load(s + 2);
        if (normalise) { v0 = (v0 - nm.mean[0]) * nm.inv_std[0]; v1 = (v1 - nm.mean[1]) * nm.inv_std[1]; v2 = (v2 - nm.mean[2]) * nm.inv_std[2]; }
        __half* d = out + i * 3;
        d[0] = __float2half_rn(v0); d[1] = __float2half_rn(v1); d[2] = __float2half_rn(v2);
    }
}

// tokens[img, 0] = cls + pos[0]; tokens[img, 1 + i] = patches[img, i] + pos[1 + i]      (8 halves per thread)
__global__ void __launch_bounds__(256) k_vit_assemble(const uint4* __restrict__ patches, const uint4* __restrict__ cls, const uint4* __restrict__ pos,
                                                      uint4* __restrict__ out, int n, int np, int cv) {
    const long long total = (long long)n * (np + 1) * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv); const long long r = i / cv; const int tok = (int)(r % (np + 1)); const long long img = r / (np + 1);
        const uint4 a = tok == 0 ? __ldg(cls + v) : __ldg(patches + (img * np + tok - 1) * cv + v);
        const uint4 b = __ldg(pos + (long long)tok * cv + v);
        uint4 o; const __half2* ah = reinterpret_cast<const __half2*>(&a); const __half2* bh = reinterpret_cast<const __half2*>(&b); __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 fa = __half22float2(ah[j]), fb = __half22float2(bh[j]); oh[j] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y); }
        out[i] = o;
    }
}

// out[r] = in[r] / max(|in[r]|, eps): one warp per row, fp32 math
__global__ void __launch_bounds__(256) k_l2norm_rows(const __half* __restrict__ in, long long ld_in, __half* __restrict__ out, long long ld_out,
                                                     long long rows, int c, float eps) {
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
        const __half* s = in + r * ld_in;
        float q = 0.f;
        for (int i = lane; i < c; i += 32) { const float v = __half2float(s[i]); q = fmaf(v, v, q); }
        q = warp_sum(q);
        const float inv = 1.0f / fmaxf(sqrtf(q), eps);
        __half* d = out + r * ld_out;
        for (int i = lane; i < c; i += 32) d[i] = __float2half_rn(__half2float(s[i]) * inv);
    }
}

// sum over a rows x cols window of (a - b)^2, fp32 inputs, double accumulation
__global__ void __launch_bounds__(256) k_sqdiff_f32(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
                                                    long long rows, int cols, double* __restrict__ out) {
    __shared__ double red[8];
    const long long total = rows * cols;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols; const int c = (int)(i - r * cols);
        const double d = (double)a[r * lda + c] - (double)b[r * ldb + c];
        acc += d * d;
    }
    const double s = block_sum_double(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}

// out[r] = <a[r], b[r]> / max(|a[r]| |b[r]|, eps): one warp per row pair
__global__ void __launch_bounds__(256) k_cosine_rows(const __half* __restrict__ a, long long lda, const __half* __restrict__ b, long long ldb,
                                                     float* __restrict__ out, long long rows, int c, float eps) {
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (blockDim.x >> 5)) {
        float ab = 0.f, aa = 0.f, bb = 0.f;
        for (int i = lane; i < c; i += 32) { const float x = __half2float(a[r * lda + i]), y = __half2float(b[r * ldb + i]); ab = fmaf(x, y, ab); aa = fmaf(x, x, aa); bb = fmaf(y, y, bb); }
        ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
        if (lane == 0) out[r] = ab / fmaxf(sqrtf(aa) * sqrtf(bb), eps);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// LPIPS (SqueezeNet 1.1) helpers.  Activations are NHWC fp16 with odd spatial sizes (255, 127, 63, 31 for a 512^2 input), which the
// TMA-im2col convolution of gemm_conv.cu does not tile; its 3x3 convolutions therefore go through an explicit im2col + fie_gemm_f16.
// ------------------------------------------------------------------------------------------------------------------------
// out[(img, oy, ox), (ky, kx, c)] = in[img, oy s + ky - pad, ox s + kx - pad, c] (zero outside), rows padded with zeros to kpad columns.
// u8_input: the network input, uint8 RGB mapped to ((v / 255 * 2 - 1) - shift[c]) / scale[c] (LPIPS ScalingLayer on a [-1, 1] image).
template <typename T>
__global__ void __launch_bounds__(256) k_im2col3x3(const T* __restrict__ in, long long ld_in, __half* __restrict__ out, int n, int h, int w, int c,
                                                   int oh, int ow, int stride, int pad, int kpad, Norm3 nm) {
    const long long total = (long long)n * oh * ow * kpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % kpad); long long t = i / kpad;
        const int ox = (int)(t % ow); t /= ow; const int oy = (int)(t % oh); const long long img = t / oh;
        float v = 0.f;
        if (col < 9 * c) {
            const int tap = col / c, ch = col - tap * c, ky = tap / 3, kx = tap - 3 * ky;
            const int y = oy * stride + ky - pad, x = ox * stride + kx - pad;
            if (y >= 0 && y < h && x >= 0 && x < w) {
                const T* s = in + ((img * h + y) * w + x) * ld_in + ch;
                if (sizeof(T) == 1) v = ((rs_load(s) * 2.0f - 1.0f) - nm.mean[ch]) * nm.inv_std[ch];
                else v = rs_load(s);
            }
        }
        out[i] = __float2half_rn(v);
    }
}

// 3x3 stride-2 max pooling, ceil_mode=True, no padding (windows are clipped at the border); 8 channels per thread
__global__ void __launch_bounds__(256) k_maxpool3s2(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int h, int w, int cv, int oh, int ow) {
    const long long total = (long long)n * oh * ow * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv); long long t = i / cv;
        const int ox = (int)(t % ow); t /= ow; const int oy = (int)(t % oh); const long long img = t / oh;
        __half2 m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __float2half2_rn(-65504.0f);
        for (int ky = 0; ky < 3; ++ky) {
            const int y = 2 * oy + ky; if (y >= h) break;
            for (int kx = 0; kx < 3; ++kx) {
                const int x = 2 * ox + kx; if (x >= w) break;
                const uint4 q = __ldg(in + ((img * h + y) * w + x) * cv + v);
                const __half2* qh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], qh[j]);
            }
        }
        uint4 o; __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh2[j] = m[j];
        out[i] = o;
    }
}

// One LPIPS layer: for every pixel unit-normalise both feature vectors over channels (x / (|x| + 1e-10)), squared difference,
// dot with the non-negative "lin" weights; spatial SUM per image into out[img] (double).  One warp per pixel.
__global__ void __launch_bounds__(256) k_lpips_layer(const __half* __restrict__ f0, const __half* __restrict__ f1, const float* __restrict__ lin,
                                                     int n, long long hw, int c, double* __restrict__ out) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, img = blockIdx.y;
    double acc = 0.0;
    for (long long px = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); px < hw; px += (long long)gridDim.x * (blockDim.x >> 5)) {
        const __half* a = f0 + ((long long)img * hw + px) * c; const __half* b = f1 + ((long long)img * hw + px) * c;
        float qa = 0.f, qb = 0.f;
        for (int i = lane; i < c; i += 32) { const float x = __half2float(a[i]), y = __half2float(b[i]); qa = fmaf(x, x, qa); qb = fmaf(y, y, qb); }
        qa = warp_sum(qa); qb = warp_sum(qb);
        const float ia = 1.0f / (sqrtf(qa) + 1e-10f), ib = 1.0f / (sqrtf(qb) + 1e-10f);
        float s = 0.f;
        for (int i = lane; i < c; i += 32) {
            const float d = __fmul_rn(__half2float(a[i]), ia) - __fmul_rn(__half2float(b[i]), ib);     // no FMA contraction: identical features give exactly 0
            s = fmaf(__ldg(lin + i), d * d, s);
        }
        acc += (double)s;                                   // every lane holds a partial: summed over the block below
    }
    const double tot = block_sum_double(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + img, tot);
}

}  // namespace fie
using namespace fie;

#define FIE_ZERO(ptr, bytes, stream, what)                                                                    \
    do { cudaError_t e_ = cudaMemsetAsync((ptr), 0, (bytes), (stream));                                       \
         if (e_ != cudaSuccess) { set_error("%s: memset: %s", (what), cudaGetErrorString(e_)); return FIE_ERR_CUDA; } } while (0)

extern "C" int fie_ssim_u8(const void* a, const void* b, int n, int h, int w, int c, int kernel_size, double sigma, double k1, double k2,
                           double* out_sum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(a && b && out_sum && n > 0 && c > 0, "fie_ssim_u8: bad args");
    FIE_REQUIRE(kernel_size >= 1 && kernel_size <= SS_MAXK && (kernel_size & 1) && sigma > 0., "fie_ssim_u8: kernel_size must be odd and <= %d", SS_MAXK);
    FIE_REQUIRE(h >= kernel_size && w >= kernel_size, "fie_ssim_u8: image smaller than the window");
    FIE_REQUIRE((long long)n * c <= 65535, "fie_ssim_u8: too many (image, channel) planes");
    SsimWeights wt;
    double tot = 0.;
    for (int i = 0; i < kernel_size; ++i) { const double d = (double)(i - (kernel_size - 1) / 2) / sigma; wt.w[i] = exp(-(d * d) / 2.0); tot += wt.w[i]; }
    for (int i = 0; i < kernel_size; ++i) wt.w[i] /= tot;
    for (int i = kernel_size; i < SS_MAXK; ++i) wt.w[i] = 0.;
    FIE_ZERO(out_sum, sizeof(double) * n, stream, "fie_ssim_u8");
    const int vh = h - kernel_size + 1, vw = w - kernel_size + 1;
    dim3 grid(ceil_div(vw, SS_TW), ceil_div(vh, SS_TH), n * c);
    FIE_REQUIRE(grid.y <= 65535, "fie_ssim_u8: image too tall");
    k_ssim_u8<<<grid, 256, 0, stream>>>((const uint8_t*)a, (const uint8_t*)b, h, w, c, kernel_size, wt, k1 * k1, k2 * k2, out_sum);
    return check_launch("fie_ssim_u8");
}

extern "C" int fie_sqdiff_u8(const void* a, const void* b, int n, long long per_image, unsigned long long* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(a && b && out && n > 0 && n <= 65535 && per_image > 0, "fie_sqdiff_u8: bad args");
    FIE_ZERO(out, sizeof(unsigned long long) * n, stream, "fie_sqdiff_u8");
    const int vec16 = (per_image % 16) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    unsigned gx = mgrid(vec16 ? per_image / 16 : per_image, 256); const unsigned cap = (unsigned)((148 * 16 + n - 1) / n); if (gx > cap) gx = cap;
    k_sqdiff_u8<<<dim3(gx, n), 256, 0, stream>>>((const uint8_t*)a, (const uint8_t*)b, per_image, vec16, out);
    return check_launch("fie_sqdiff_u8");
}

extern "C" int fie_resample_f32(const void* in, int in_is_u8, void* out, void* tmp, int n, int h, int w, int oh, int ow,
                                const int* bounds_x, const float* coeff_x, int ksize_x, const int* bounds_y, const float* coeff_y, int ksize_y,
                                const float* mean3, const float* std3, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "fie_resample_f32: bad shape");
    const bool need_h = ow != w, need_v = oh != h;
    FIE_REQUIRE(!need_h || (bounds_x && coeff_x && ksize_x > 0), "fie_resample_f32: horizontal tables missing");
    FIE_REQUIRE(!need_v || (bounds_y && coeff_y && ksize_y > 0), "fie_resample_f32: vertical tables missing");
    FIE_REQUIRE(!need_h || tmp, "fie_resample_f32: the horizontal pass needs the fp32 [n,h,ow,3] temporary");
    Norm3 nm;
    for (int i = 0; i < 3; ++i) {
        nm.mean[i] = mean3 ? mean3[i] : 0.f;
        const float s = std3 ? std3[i] : 1.f;
        FIE_REQUIRE(s != 0.f, "fie_resample_f32: zero std");
        nm.inv_std[i] = 1.0f / s;
    }
    const int* by = need_v ? bounds_y : nullptr;
    if (need_h) {
        if (in_is_u8) k_resample_f32_h<uint8_t><<<mgrid((long long)n * h * ow, 256), 256, 0, stream>>>((const uint8_t*)in, (float*)tmp, n, h, w, ow, bounds_x, coeff_x, ksize_x);
        else k_resample_f32_h<float><<<mgrid((long long)n * h * ow, 256), 256, 0, stream>>>((const float*)in, (float*)tmp, n, h, w, ow, bounds_x, coeff_x, ksize_x);
        k_resample_f32_v<float><<<mgrid((long long)n * oh * ow, 256), 256, 0, stream>>>((const float*)tmp, (float*)out, n, h, ow, oh, by, coeff_y, ksize_y, nm);
    } else if (in_is_u8) {
        k_resample_f32_v<uint8_t><<<mgrid((long long)n * oh * ow, 256), 256, 0, stream>>>((const uint8_t*)in, (float*)out, n, h, ow, oh, by, coeff_y, ksize_y, nm);
    } else {
        k_resample_f32_v<float><<<mgrid((long long)n * oh * ow, 256), 256, 0, stream>>>((const float*)in, (float*)out, n, h, ow, oh, by, coeff_y, ksize_y, nm);
    }
    return check_launch("fie_resample_f32");
}

extern "C" int fie_patchify_f16(const void* in, int in_is_u8, void* out, int n, int h, int w, int patch, const float* mean3, const float* std3, void* stream_) {
    FIE_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && patch > 0 && (h % patch) == 0 && (w % patch) == 0, "fie_patchify_f16: h and w must be multiples of the patch size");
    Norm3 nm;
    const int normalise = mean3 && std3;
    for (int i = 0; i < 3; ++i) { nm.mean[i] = normalise ? mean3[i] : 0.f; nm.inv_std[i] = normalise ? 1.0f / std3[i] : 1.f; }
    const long long total = (long long)n * h * w;
    if (in_is_u8) k_patchify<uint8_t><<<mgrid(total, 256), 256, 0, (cudaStream_t)stream_>>>((const uint8_t*)in, (__half*)out, n, h, w, patch, nm, normalise);
    else k_patchify<float><<<mgrid(total, 256), 256, 0, (cudaStream_t)stream_>>>((const float*)in, (__half*)out, n, h, w, patch, nm, normalise);
    return check_launch("fie_patchify_f16");
}

extern "C" int fie_vit_assemble_f16(const void* patches, const void* cls, const void* pos, void* out, int n, int n_patches, int c, void* stream_) {
    FIE_REQUIRE(patches && cls && pos && out && n > 0 && n_patches > 0 && c > 0 && (c % 8) == 0, "fie_vit_assemble_f16: bad args (c must be a multiple of 8)");
    k_vit_assemble<<<mgrid((long long)n * (n_patches + 1) * (c / 8), 256), 256, 0, (cudaStream_t)stream_>>>((const uint4*)patches, (const uint4*)cls, (const uint4*)pos,
                                                                                                        (uint4*)out, n, n_patches, c / 8);
    return check_launch("fie_vit_assemble_f16");
}

extern "C" int fie_l2norm_rows_f16(const void* in, long long ld_in, void* out, long long ld_out, long long rows, int c, float eps, void* stream_) {
    FIE_REQUIRE(in && out && rows > 0 && c > 0 && ld_in >= c && ld_out >= c && eps > 0.f, "fie_l2norm_rows_f16: bad args");
    k_l2norm_rows<<<mgrid(rows, 8), 256, 0, (cudaStream_t)stream_>>>((const __half*)in, ld_in, (__half*)out, ld_out, rows, c, eps);
    return check_launch("fie_l2norm_rows_f16");
}

extern "C" int fie_sqdiff_f32(const void* a, long long lda, const void* b, long long ldb, long long rows, int cols, double* out_sum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(a && b && out_sum && rows > 0 && cols > 0 && lda >= cols && ldb >= cols, "fie_sqdiff_f32: bad args");
    FIE_ZERO(out_sum, sizeof(double), stream, "fie_sqdiff_f32");
    k_sqdiff_f32<<<mgrid(rows * cols, 256), 256, 0, stream>>>((const float*)a, lda, (const float*)b, ldb, rows, cols, out_sum);
    return check_launch("fie_sqdiff_f32");
}

extern "C" int fie_cosine_rows_f16(const void* a, long long lda, const void* b, long long ldb, float* out, long long rows, int c, float eps, void* stream_) {
    FIE_REQUIRE(a && b && out && rows > 0 && c > 0 && lda >= c && ldb >= c && eps > 0.f, "fie_cosine_rows_f16: bad args");
    k_cosine_rows<<<mgrid(rows, 8), 256, 0, (cudaStream_t)stream_>>>((const __half*)a, lda, (const __half*)b, ldb, out, rows, c, eps);
    return check_launch("fie_cosine_rows_f16");
}

extern "C" int fie_im2col3x3_f16(const void* in, int in_is_u8, long long ld_in, void* out, int n, int h, int w, int c, int stride, int pad, int kpad,
                                 const float* shift3, const float* scale3, void* stream_) {
    FIE_REQUIRE(in && out && n > 0 && h > 0 && w > 0 && c > 0 && stride >= 1 && pad >= 0 && pad <= 1, "fie_im2col3x3_f16: bad args");
    FIE_REQUIRE(kpad >= 9 * c && ld_in >= c, "fie_im2col3x3_f16: kpad must cover 9 * c columns");
    FIE_REQUIRE(h + 2 * pad >= 3 && w + 2 * pad >= 3, "fie_im2col3x3_f16: input smaller than the window");
    FIE_REQUIRE(!in_is_u8 || (c == 3 && shift3 && scale3), "fie_im2col3x3_f16: the uint8 form is the RGB network input (needs shift / scale)");
    const int oh = (h + 2 * pad - 3) / stride + 1, ow = (w + 2 * pad - 3) / stride + 1;
    Norm3 nm;
    for (int i = 0; i < 3; ++i) { nm.mean[i] = shift3 ? shift3[i] : 0.f; nm.inv_std[i] = scale3 ? 1.0f / scale3[i] : 1.f; }
    const long long total = (long long)n * oh * ow * kpad;
    if (in_is_u8) k_im2col3x3<uint8_t><<<mgrid(total, 256), 256, 0, (cudaStream_t)stream_>>>((const uint8_t*)in, ld_in, (__half*)out, n, h, w, c, oh, ow, stride, pad, kpad, nm);
    else k_im2col3x3<__half><<<mgrid(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __half*)in, ld_in, (__half*)out, n, h, w, c, oh, ow, stride, pad, kpad, nm);
    return check_launch("fie_im2col3x3_f16");
}

extern "C" int fie_maxpool3s2_ceil_f16(const void* in, void* out, int n, int h, int w, int c, void* stream_) {
    FIE_REQUIRE(in && out && n > 0 && h >= 3 && w >= 3 && c > 0 && (c % 8) == 0, "fie_maxpool3s2_ceil_f16: bad args (c must be a multiple of 8)");
    const int oh = (h - 3 + 1) / 2 + 1, ow = (w - 3 + 1) / 2 + 1;                 // ceil((h - 3) / 2) + 1; the last window always starts inside
    k_maxpool3s2<<<mgrid((long long)n * oh * ow * (c / 8), 256), 256, 0, (cudaStream_t)stream_>>>((const uint4*)in, (uint4*)out, n, h, w, c / 8, oh, ow);
    return check_launch("fie_maxpool3s2_ceil_f16");
}

extern "C" int fie_lpips_layer_f16(const void* f0, const void* f1, const float* lin, int n, long long hw, int c, double* out_sum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(f0 && f1 && lin && out_sum && n > 0 && n <= 65535 && hw > 0 && c > 0, "fie_lpips_layer_f16: bad args");
    FIE_ZERO(out_sum, sizeof(double) * n, stream, "fie_lpips_layer_f16");
    unsigned gx = mgrid(hw, 8); const unsigned cap = (unsigned)((148 * 8 + n - 1) / n); if (gx > cap) gx = cap;
    k_lpips_layer<<<dim3(gx, n), 256, 0, stream>>>((const __half*)f0, (const __half*)f1, lin, n, hw, c, out_sum);
    return check_launch("fie_lpips_layer_f16");
}

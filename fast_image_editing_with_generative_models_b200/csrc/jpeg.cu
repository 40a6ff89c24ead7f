// Baseline JPEG encoder on the GPU: replaces `edited.save(output_path)` of the reference's callers
// (run_batch.py:224, run_single_image.py:114 = PIL -> libjpeg(-turbo): baseline sequential DCT, quality 75, 4:2:0, Annex K Huffman
// tables, JFIF header).  The output is BYTE-IDENTICAL to Pillow's file for the same pixels (tests/test_gpu_jpeg.py), so the sweep can
// write its *.jpg files from compressed bytes that left the GPU already encoded: ~100-300 KB instead of 3 MB per image over PCIe and
// no host-side encode (at 165 edits/s per box the host JPEG path is otherwise the next cap).
//
// Integer only.  Per image:
//   k_jpeg_dct    one 64-thread group per 16x16 MCU: RGB -> YCbCr (16-bit fixed point, libjpeg jccolor.c), h2v2 chroma downsampling with
//                 the alternating 1,2 bias (jcsample.c), edge replication in libjpeg's two stages (jcprepct.c), level shift, the "islow"
//                 integer forward DCT (jfdctint.c), quantisation (round-half-up of the magnitude by 8 q, jcdctmgr.c), zig-zag, and the
//                 dummy-block rule for luma blocks outside the image (jccoefct.c) -> int16 coefficients [mcu][6][64]
//   k_jpeg_count  one thread per block: number of Huffman bits of the block (DC difference against the previous block of the component)
//   k_jpeg_scan   one CTA per image: exclusive prefix sum of the bit counts -> bit offset of every block, total bits
//   k_jpeg_emit   one thread per block: writes the block's codes at its bit offset (atomicOr into a zeroed word stream, MSB first)
//   k_jpeg_ffcount / k_jpeg_scan / k_jpeg_stuff   byte stuffing (0xFF -> 0xFF 0x00): count, scan, scatter; pads the last byte with
//                 1-bits, prepends the header (built on the host once per size / quality) and appends EOI
#include "fie_common.cuh"
#include <string.h>

namespace fie {

__constant__ uint16_t c_huff_code[4][256];     // 0 DC luma, 1 AC luma, 2 DC chroma, 3 AC chroma
__constant__ uint8_t  c_huff_size[4][256];
__constant__ uint8_t  c_zigzag[64];            // natural index of zig-zag position k

struct JpegQ { uint16_t q[2][64]; };           // quantisation tables (natural order): luma, chroma

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// One 8-point pass of jpeg_fdct_islow; FIRST = row pass (results scaled up by 2^PASS1_BITS), else the column pass.
template <bool FIRST>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
    constexpr int CB = 13, P1 = 2;
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int SH = FIRST ? CB - P1 : CB + P1;
    if (FIRST) { d[0] = (t10 + t11) << P1; d[4] = (t10 - t11) << P1; }
    else { d[0] = descale(t10 + t11, P1); d[4] = descale(t10 - t11, P1); }
    int z1 = (t12 + t13) * 4433;
    d[2] = descale(z1 + t13 * 6270, SH);
    d[6] = descale(z1 + t12 * (-15137), SH);
    z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int u4 = t4 * 2446, u5 = t5 * 16819, u6 = t6 * 25172, u7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 = z3 * (-16069) + z5; z4 = z4 * (-3196) + z5;
    d[7] = descale(u4 + z1 + z3, SH); d[5] = descale(u5 + z2 + z4, SH);
    d[3] = descale(u6 + z2 + z3, SH); d[1] = descale(u7 + z1 + z4, SH);
}

constexpr int MCU_PER_CTA = 4;
__global__ void __launch_bounds__(64 * MCU_PER_CTA) k_jpeg_dct(const uint8_t* __restrict__ rgb, int16_t* __restrict__ coef, int H, int W, int mcu_cols,
                                                               int mcus_per_image, JpegQ Q) {
    __shared__ int s_blk[MCU_PER_CTA][6][64];
    const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int img = blockIdx.y;
    const int mcu = blockIdx.x * MCU_PER_CTA + g;
    const bool live = mcu < mcus_per_image;
    const int my = live ? mcu / mcu_cols : 0, mx = live ? mcu % mcu_cols : 0;
    const uint8_t* src = rgb + (size_t)img * H * W * 3;
    int (*blk)[64] = s_blk[g];
    // ---- colour conversion: 256 luma samples (4 per thread) and 64 chroma samples (1 per thread) of the MCU ----
    if (live) {
        const int hc = (H + 1) >> 1;                       // real chroma rows: later rows replicate the last DOWNSAMPLED row (jcprepct.c)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int p = t + 64 * i, ly = p >> 4, lx = p & 15;
            const int y = min(my * 16 + ly, H - 1), x = min(mx * 16 + lx, W - 1);
            const uint8_t* px = src + ((size_t)y * W + x) * 3;
            const int r = px[0], gg = px[1], b = px[2];
            const int Y = (19595 * r + 38470 * gg + 7471 * b + 32768) >> 16;
            blk[(ly >> 3) * 2 + (lx >> 3)][(ly & 7) * 8 + (lx & 7)] = Y - 128;
        }
        {
            const int cy = t >> 3, cx = t & 7;
            const int rc = min(my * 8 + cy, hc - 1);
            const int y0 = min(2 * rc, H - 1), y1 = min(2 * rc + 1, H - 1);
            const int gx = mx * 8 + cx;
            const int x0 = min(2 * gx, W - 1), x1 = min(2 * gx + 1, W - 1);
            int sb = 0, sr = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint8_t* px = src + ((size_t)((k & 2) ? y1 : y0) * W + ((k & 1) ? x1 : x0)) * 3;
                const int r = px[0], gg = px[1], b = px[2];
                sb += (-11059 * r - 21709 * gg + 32768 * b + 8388608 + 32767) >> 16;
                sr += (32768 * r - 27439 * gg - 5329 * b + 8388608 + 32767) >> 16;
            }
            const int bias = (gx & 1) ? 2 : 1;
            blk[4][t] = ((sb + bias) >> 2) - 128;
            blk[5][t] = ((sr + bias) >> 2) - 128;
        }
    }
    __syncthreads();
    // ---- forward DCT: 48 row passes, then 48 column passes (one 8-point transform per thread) ----
    if (live && t < 48) {
        int d[8]; int* row = &blk[t >> 3][(t & 7) * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = row[i];
        fdct8<true>(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) row[i] = d[i];
    }
    __syncthreads();
    if (live && t < 48) {
        int d[8]; int* col = &blk[t >> 3][t & 7];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = col[i * 8];
        fdct8<false>(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) col[i * 8] = d[i];
    }
    __syncthreads();
    // ---- quantise + zig-zag: 384 coefficients per MCU, 6 per thread ----
    if (live) {
        const int by = (H + 7) >> 3, bx = (W + 7) >> 3;          // real luma blocks; blocks beyond are dummies (jccoefct.c)
        int16_t* out = coef + ((size_t)img * mcus_per_image + mcu) * 384;
        int dc[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {                            // DC of the luma blocks after the dummy rule (every thread computes them)
            const int q = (int)Q.q[0][0] << 3; const int v = blk[b][0]; const int a = abs(v);
            const int qv = (a + (q >> 1)) / q;
            dc[b] = v < 0 ? -qv : qv;
        }
        const bool dum_r = (mx * 2 + 1) >= bx, dum_b = (my * 2 + 1) >= by;
        if (dum_r) dc[1] = dc[0];
        if (dum_b) { dc[2] = dc[1]; dc[3] = dc[1]; } else if (dum_r) dc[3] = dc[2];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int e = t + 64 * i, b = e >> 6, k = e & 63;
            const int nat = c_zigzag[k];
            const int q = (int)Q.q[b < 4 ? 0 : 1][nat] << 3;
            const int v = blk[b][nat]; const int a = abs(v);
            int qv = (a + (q >> 1)) / q;
            qv = v < 0 ? -qv : qv;
            if (b < 4) {
                const bool dummy = (b == 1 && dum_r) || (b == 2 && dum_b) || (b == 3 && (dum_r || dum_b));
                if (dummy) qv = 0;
                if (k == 0) qv = dc[b];
            }
            out[e] = (int16_t)qv;
        }
    }
}

__device__ __forceinline__ int bit_length(int v) { return 32 - __clz(v); }          // v >= 0

// Walks one block's zig-zag coefficients; EMIT = false: returns the bit count, EMIT = true: writes the codes at bit offset `pos`.
template <bool EMIT>
__device__ __forceinline__ unsigned block_bits(const int16_t* __restrict__ c, int pred, bool chroma, uint32_t* words, unsigned long long pos) {
    const int dt = chroma ? 2 : 0, at = dt + 1;
    unsigned nbits = 0;
    unsigned long long acc = 0; int accn = 0;                      // EMIT: pending bits (MSB-aligned at bit accn-1 .. 0)
    auto put = [&](unsigned code, int size) {
        nbits += size;
        if (EMIT) {
            acc = (acc << size) | (code & ((1u << size) - 1u));
            accn += size;
            if (accn >= 32) {                                      // flush one full 32-bit group at the current position
                const uint32_t w = (uint32_t)(acc >> (accn - 32));
                const unsigned sh = (unsigned)(pos & 31);
                atomicOr(&words[pos >> 5], sh ? (w >> sh) : w);
                if (sh) atomicOr(&words[(pos >> 5) + 1], w << (32 - sh));
                pos += 32; accn -= 32;
            }
        }
    };
    // 16-byte vector loads of the 64 coefficients
    __align__(16) int16_t v[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(&v[8 * i]) = __ldg(reinterpret_cast<const uint4*>(c) + i);
    const int diff = (int)v[0] - pred;
    int cat = bit_length(abs(diff));
    put(c_huff_code[dt][cat], c_huff_size[dt][cat]);
    if (cat) put((unsigned)(diff >= 0 ? diff : diff - 1), cat);
    int last = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) if (v[k] != 0) last = k;
    int run = 0;
    for (int k = 1; k <= last; ++k) {
        const int x = v[k];
        if (x == 0) { ++run; continue; }
        while (run > 15) { put(c_huff_code[at][0xF0], c_huff_size[at][0xF0]); run -= 16; }
        cat = bit_length(abs(x));
        const int sym = (run << 4) | cat;
        put(c_huff_code[at][sym], c_huff_size[at][sym]);
        put((unsigned)(x >= 0 ? x : x - 1), cat);
        run = 0;
    }
    if (last < 63) put(c_huff_code[at][0], c_huff_size[at][0]);
    if (EMIT && accn > 0) {                                        // tail: fewer than 32 pending bits
        const uint32_t w = (uint32_t)(acc << (32 - accn));
        const unsigned sh = (unsigned)(pos & 31);
        atomicOr(&words[pos >> 5], sh ? (w >> sh) : w);
        if (sh && accn > (int)(32 - sh)) atomicOr(&words[(pos >> 5) + 1], w << (32 - sh));
    }
    return nbits;
}

__device__ __forceinline__ int dc_pred(const int16_t* coef_img, int blk) {
    // scan order: MCU m = blk / 6, block b = blk % 6 (Y00 Y01 Y10 Y11 Cb Cr); predecessor = previous block of the same component
    const int m = blk / 6, b = blk - m * 6;
    if (b >= 4) return m == 0 ? 0 : coef_img[(size_t)(m - 1) * 384 + b * 64];
    if (b > 0) return coef_img[(size_t)m * 384 + (b - 1) * 64];
    return m == 0 ? 0 : coef_img[(size_t)(m - 1) * 384 + 3 * 64];
}

__global__ void __launch_bounds__(128) k_jpeg_count(const int16_t* __restrict__ coef, uint32_t* __restrict__ bits, int blocks_per_image) {
    const int img = blockIdx.y, blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= blocks_per_image) return;
    const int16_t* ci = coef + (size_t)img * blocks_per_image * 64;
    bits[(size_t)img * blocks_per_image + blk] = block_bits<false>(ci + (size_t)blk * 64, dc_pred(ci, blk), (blk % 6) >= 4, nullptr, 0);
}

// Exclusive prefix sum of `n` 32-bit counts per image into 64-bit offsets (one 1024-thread CTA per image); totals[img] = sum.
__global__ void __launch_bounds__(1024) k_jpeg_scan(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, unsigned long long* __restrict__ totals, int n) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* a = in + (size_t)img * n;
    unsigned long long* o = out + (size_t)img * n;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const unsigned long long v = i < n ? a[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long s = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += y; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const unsigned long long before = carry + (wid ? warp_sums[wid - 1] : 0ull) + (x - v);
        if (i < n) o[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) totals[img] = carry;
}

__global__ void __launch_bounds__(128) k_jpeg_emit(const int16_t* __restrict__ coef, const unsigned long long* __restrict__ offs, uint32_t* __restrict__ words,
                                                   int blocks_per_image, size_t words_per_image) {
    const int img = blockIdx.y, blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= blocks_per_image) return;
    const int16_t* ci = coef + (size_t)img * blocks_per_image * 64;
    block_bits<true>(ci + (size_t)blk * 64, dc_pred(ci, blk), (blk % 6) >= 4, words + (size_t)img * words_per_image, offs[(size_t)img * blocks_per_image + blk]);
}

// Raw scan byte k of an image (MSB-first words); the last byte is padded with 1-bits (libjpeg flush_bits).
__device__ __forceinline__ unsigned raw_byte(const uint32_t* w, unsigned long long k, unsigned long long total_bits) {
    unsigned b = (w[k >> 2] >> (24 - 8 * (unsigned)(k & 3))) & 0xFFu;
    if ((k + 1) * 8 > total_bits) b |= 0xFFu >> (unsigned)(total_bits - k * 8);
    return b;
}

constexpr int STUFF_CHUNK = 32;       // raw bytes per thread in the stuffing passes
__global__ void __launch_bounds__(256) k_jpeg_ffcount(const uint32_t* __restrict__ words, const unsigned long long* __restrict__ total_bits, uint32_t* __restrict__ counts,
                                                      size_t words_per_image, int chunks_per_image) {
    const int img = blockIdx.y, ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= chunks_per_image) return;
    const unsigned long long tb = total_bits[img], nbytes = (tb + 7) >> 3;
    const uint32_t* w = words + (size_t)img * words_per_image;
    unsigned cnt = 0;
    const unsigned long long k0 = (unsigned long long)ch * STUFF_CHUNK;
    for (int i = 0; i < STUFF_CHUNK; ++i) { const unsigned long long k = k0 + i; if (k < nbytes) cnt += raw_byte(w, k, tb) == 0xFFu; }
    counts[(size_t)img * chunks_per_image + ch] = cnt;
}

__global__ void __launch_bounds__(256) k_jpeg_stuff(const uint32_t* __restrict__ words, const unsigned long long* __restrict__ total_bits,
                                                    const unsigned long long* __restrict__ ff_before, const unsigned long long* __restrict__ ff_total,
                                                    const uint8_t* __restrict__ header, int header_bytes, uint8_t* __restrict__ out, size_t out_stride,
                                                    int* __restrict__ out_sizes, size_t words_per_image, int chunks_per_image) {
    const int img = blockIdx.y, ch = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long tb = total_bits[img], nbytes = (tb + 7) >> 3;
    uint8_t* o = out + (size_t)img * out_stride;
    if (blockIdx.x == 0) {                                         // header, EOI and the file size
        for (int i = threadIdx.x; i < header_bytes; i += blockDim.x) o[i] = header[i];
        if (threadIdx.x == 0) {
            const unsigned long long end = (unsigned long long)header_bytes + nbytes + ff_total[img];
            o[end] = 0xFF; o[end + 1] = 0xD9;
            out_sizes[img] = (int)(end + 2);
        }
    }
    if (ch >= chunks_per_image) return;
    const uint32_t* w = words + (size_t)img * words_per_image;
    const unsigned long long k0 = (unsigned long long)ch * STUFF_CHUNK;
    unsigned long long dst = (unsigned long long)header_bytes + k0 + ff_before[(size_t)img * chunks_per_image + ch];
    for (int i = 0; i < STUFF_CHUNK; ++i) {
        const unsigned long long k = k0 + i;
        if (k >= nbytes) break;
        const unsigned b = raw_byte(w, k, tb);
        o[dst++] = (uint8_t)b;
        if (b == 0xFFu) o[dst++] = 0;
    }
}

}  // namespace fie
using namespace fie;

// ---- host: tables ----
static const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t kLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                   18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
static const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

static void derive_huff(const uint8_t* bits, const uint8_t* vals, uint16_t* code, uint8_t* size) {
    memset(code, 0, 256 * sizeof(uint16_t)); memset(size, 0, 256);
    unsigned c = 0; int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) { code[vals[k]] = (uint16_t)c; size[vals[k]] = (uint8_t)len; ++c; ++k; }
        c <<= 1;
    }
}

static void quality_tables(int quality, uint16_t (*q)[64]) {
    if (quality < 1) quality = 1; if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            long v = ((long)(t ? kChromaQ[i] : kLumaQ[i]) * scale + 50) / 100;
            q[t][i] = (uint16_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
        }
}

static int upload_tables_once() {
    static bool done[kMaxDevices] = {false};
    bool& d = done[current_device()];
    if (d) return FIE_OK;
    uint16_t code[4][256]; uint8_t size[4][256];
    derive_huff(kDcLumaBits, kDcVals, code[0], size[0]); derive_huff(kAcLumaBits, kAcLumaVals, code[1], size[1]);
    derive_huff(kDcChromaBits, kDcVals, code[2], size[2]); derive_huff(kAcChromaBits, kAcChromaVals, code[3], size[3]);
    cudaError_t e = cudaMemcpyToSymbol(c_huff_code, code, sizeof(code));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_huff_size, size, sizeof(size));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_zigzag, kZigzag, sizeof(kZigzag));
    if (e != cudaSuccess) { set_error("fie_jpeg: table upload: %s", cudaGetErrorString(e)); return FIE_ERR_CUDA; }
    d = true;
    return FIE_OK;
}

extern "C" int fie_jpeg_header_bytes(void) { return 2 + 18 + 2 * 69 + 19 + 2 * 33 + 2 * 183 + 14; }     // 623

// The segments Pillow / libjpeg write before the scan: SOI, JFIF APP0, 2 x DQT, SOF0 (4:2:0), 4 x DHT, SOS.
extern "C" int fie_jpeg_write_header(uint8_t* dst, int h, int w, int quality) {
    FIE_REQUIRE(dst && h > 0 && w > 0 && h < 65536 && w < 65536, "fie_jpeg_write_header: bad arguments");
    uint16_t q[2][64]; quality_tables(quality, q);
    uint8_t* p = dst;
    auto seg = [&](uint8_t marker, const uint8_t* payload, int n) { *p++ = 0xFF; *p++ = marker; *p++ = (uint8_t)((n + 2) >> 8); *p++ = (uint8_t)(n + 2); memcpy(p, payload, n); p += n; };
    *p++ = 0xFF; *p++ = 0xD8;
    const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    seg(0xE0, app0, 14);
    for (int t = 0; t < 2; ++t) { uint8_t d[65]; d[0] = (uint8_t)t; for (int k = 0; k < 64; ++k) d[1 + k] = (uint8_t)q[t][kZigzag[k]]; seg(0xDB, d, 65); }
    const uint8_t sof[15] = {8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
    seg(0xC0, sof, 15);
    struct { uint8_t id; const uint8_t* bits; const uint8_t* vals; int n; } tabs[4] = {{0x00, kDcLumaBits, kDcVals, 12}, {0x10, kAcLumaBits, kAcLumaVals, 162},
                                                                                     {0x01, kDcChromaBits, kDcVals, 12}, {0x11, kAcChromaBits, kAcChromaVals, 162}};
    for (auto& t : tabs) { uint8_t d[1 + 16 + 162]; d[0] = t.id; memcpy(d + 1, t.bits, 16); memcpy(d + 17, t.vals, t.n); seg(0xC4, d, 17 + t.n); }
    const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    seg(0xDA, sos, 10);
    return (int)(p - dst);
}

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }
struct JpegLayout { int mcu_cols, mcu_rows, mcus, blocks, chunks; size_t words_per_image, coef, bits, offs, totals, words, ffcnt, ffoff, fftot, header, total; };
static JpegLayout jpeg_layout(int n, int h, int w) {
    JpegLayout L;
    L.mcu_cols = (w + 15) / 16; L.mcu_rows = (h + 15) / 16; L.mcus = L.mcu_cols * L.mcu_rows; L.blocks = L.mcus * 6;
    const size_t raw_cap = (size_t)L.blocks * 212;                 // worst case 1665 bits per block
    L.words_per_image = (raw_cap + 3) / 4 + 2;
    L.chunks = (int)((raw_cap + STUFF_CHUNK - 1) / STUFF_CHUNK);
    size_t o = 0;
    L.coef = o;   o += al256((size_t)n * L.blocks * 64 * sizeof(int16_t));
    L.bits = o;   o += al256((size_t)n * L.blocks * sizeof(uint32_t));
    L.offs = o;   o += al256((size_t)n * L.blocks * sizeof(unsigned long long));
    L.totals = o; o += al256((size_t)n * sizeof(unsigned long long));
    L.words = o;  o += al256((size_t)n * L.words_per_image * sizeof(uint32_t));
    L.ffcnt = o;  o += al256((size_t)n * L.chunks * sizeof(uint32_t));
    L.ffoff = o;  o += al256((size_t)n * L.chunks * sizeof(unsigned long long));
    L.fftot = o;  o += al256((size_t)n * sizeof(unsigned long long));
    L.header = o; o += al256(1024);
    L.total = o;
    return L;
}

extern "C" size_t fie_jpeg_workspace_bytes(int n, int h, int w) { return (n <= 0 || h <= 0 || w <= 0) ? 0 : jpeg_layout(n, h, w).total; }
// Bytes to reserve per image in `out` (worst case of a baseline JPEG: every raw byte stuffed); typical files are 3-10 % of this.
extern "C" size_t fie_jpeg_max_bytes(int h, int w) {
    if (h <= 0 || w <= 0) return 0;
    const JpegLayout L = jpeg_layout(1, h, w);
    return al256((size_t)fie_jpeg_header_bytes() + 2 * (size_t)L.blocks * 212 + 2);
}

// rgb: uint8 [n,h,w,3] (device) -> out: n JPEG files, image i at out + i * out_stride, its length in out_sizes[i] (device int32).
extern "C" int fie_jpeg_encode_u8(const void* rgb, int n, int h, int w, int quality, void* out, size_t out_stride, int* out_sizes,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FIE_REQUIRE(n >= 0 && h > 0 && w > 0 && h < 65536 && w < 65536 && n <= 65535, "fie_jpeg_encode_u8: bad shape n=%d h=%d w=%d", n, h, w);
    if (n == 0) return FIE_OK;
    FIE_REQUIRE(rgb && out && out_sizes && workspace, "fie_jpeg_encode_u8: null pointer");
    FIE_REQUIRE(out_stride >= fie_jpeg_max_bytes(h, w), "fie_jpeg_encode_u8: out_stride < fie_jpeg_max_bytes(h, w)");
    const JpegLayout L = jpeg_layout(n, h, w);
    FIE_REQUIRE(workspace_bytes >= L.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "fie_jpeg_encode_u8: workspace too small or not 256-byte aligned");
    int rc = upload_tables_once();
    if (rc) return rc;
    uint8_t* ws = (uint8_t*)workspace;
    int16_t* coef = (int16_t*)(ws + L.coef);
    uint32_t* bits = (uint32_t*)(ws + L.bits);
    unsigned long long* offs = (unsigned long long*)(ws + L.offs);
    unsigned long long* totals = (unsigned long long*)(ws + L.totals);
    uint32_t* words = (uint32_t*)(ws + L.words);
    uint32_t* ffcnt = (uint32_t*)(ws + L.ffcnt);
    unsigned long long* ffoff = (unsigned long long*)(ws + L.ffoff);
    unsigned long long* fftot = (unsigned long long*)(ws + L.fftot);
    uint8_t* dhead = ws + L.header;
    uint8_t hbuf[1024];
    const int hb = fie_jpeg_write_header(hbuf, h, w, quality);
    FIE_REQUIRE(hb == fie_jpeg_header_bytes(), "fie_jpeg_encode_u8: header size mismatch");
    cudaError_t e = cudaMemcpyAsync(dhead, hbuf, hb, cudaMemcpyHostToDevice, stream);      // 623 bytes from pageable memory: staged by the driver before it returns
    if (e == cudaSuccess) e = cudaMemsetAsync(words, 0, (size_t)n * L.words_per_image * sizeof(uint32_t), stream);
    if (e != cudaSuccess) { set_error("fie_jpeg_encode_u8: %s", cudaGetErrorString(e)); return FIE_ERR_CUDA; }
    JpegQ Q; quality_tables(quality, Q.q);
    k_jpeg_dct<<<dim3(ceil_div(L.mcus, MCU_PER_CTA), n), 64 * MCU_PER_CTA, 0, stream>>>((const uint8_t*)rgb, coef, h, w, L.mcu_cols, L.mcus, Q);
    k_jpeg_count<<<dim3(ceil_div(L.blocks, 128), n), 128, 0, stream>>>(coef, bits, L.blocks);
    k_jpeg_scan<<<n, 1024, 0, stream>>>(bits, offs, totals, L.blocks);
    k_jpeg_emit<<<dim3(ceil_div(L.blocks, 128), n), 128, 0, stream>>>(coef, offs, words, L.blocks, L.words_per_image);
    k_jpeg_ffcount<<<dim3(ceil_div(L.chunks, 256), n), 256, 0, stream>>>(words, totals, ffcnt, L.words_per_image, L.chunks);
    k_jpeg_scan<<<n, 1024, 0, stream>>>(ffcnt, ffoff, fftot, L.chunks);
    k_jpeg_stuff<<<dim3(ceil_div(L.chunks, 256), n), 256, 0, stream>>>(words, totals, ffoff, fftot, dhead, hb, (uint8_t*)out, out_stride, out_sizes,
                                                                      L.words_per_image, L.chunks);
    return check_launch("fie_jpeg_encode_u8");
}

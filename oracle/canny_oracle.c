/* TEST INFRASTRUCTURE — plain-C restatement of the reference's Canny stage (CPU baseline + checker).
 *
 * Follows the reference call site FastEditor.preprocess_image (reference src/pipeline.py:183-210):
 *   gray  = cv2.cvtColor(image_np, cv2.COLOR_RGB2GRAY)      (:200)
 *   edges = cv2.Canny(gray, low_threshold, high_threshold)  (:205)   aperture 3, L1 gradient, no blur
 *   edges_rgb = np.stack([edges]*3, axis=2)                 (:208)
 * OpenCV is an un-vendored third-party dependency (opencv-python>=4.8, requirements.txt:12); its
 * published algorithm is restated: 15-bit fixed-point gray, 3x3 Sobel with BORDER_REPLICATE,
 * L1 magnitude (0 outside the image), integer-tangent NMS (TG22 = 13573, shift 15),
 * double threshold (m > low / m > high), 8-connected hysteresis.
 * Parity status: PINNED against cv2 4.13.0 (tests/test_canny_oracle.py, tests/golden/canny_*.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TG22 13573

/* rgb: [n][h][w][3] u8 (or gray [n][h][w] when channels==1); out: [n][h][w] u8 (0/255).
 * replicate3 != 0 writes [n][h][w][3]. Returns 0 on success. */
int fie_oracle_canny_u8(const uint8_t* img, uint8_t* out, int n, int h, int w, int channels,
                        int low, int high, int replicate3)
{
    if (n < 0 || h <= 0 || w <= 0 || (channels != 1 && channels != 3)) return 1;
    if (low > high) { int t = low; low = high; high = t; }
    const size_t hw = (size_t)h * w;
    uint8_t* gray = (uint8_t*)malloc(hw);
    int32_t* mag = (int32_t*)malloc(hw * sizeof(int32_t));
    int16_t* gx = (int16_t*)malloc(hw * sizeof(int16_t));
    int16_t* gy = (int16_t*)malloc(hw * sizeof(int16_t));
    uint8_t* st = (uint8_t*)malloc(hw);           /* 0 none, 1 candidate, 2 edge */
    int32_t* stack = (int32_t*)malloc(hw * sizeof(int32_t));
    if (!gray || !mag || !gx || !gy || !st || !stack) return 2;
    for (int im = 0; im < n; ++im) {
        const uint8_t* src = img + (size_t)im * hw * channels;
        if (channels == 3) {
            for (size_t i = 0; i < hw; ++i)
                gray[i] = (uint8_t)((9798 * src[3 * i] + 19235 * src[3 * i + 1] + 3735 * src[3 * i + 2] + 16384) >> 15);
        } else memcpy(gray, src, hw);
        for (int y = 0; y < h; ++y) {
            const uint8_t* r0 = gray + (size_t)(y > 0 ? y - 1 : 0) * w;
            const uint8_t* r1 = gray + (size_t)y * w;
            const uint8_t* r2 = gray + (size_t)(y < h - 1 ? y + 1 : h - 1) * w;
            for (int x = 0; x < w; ++x) {
                int xl = x > 0 ? x - 1 : 0, xr = x < w - 1 ? x + 1 : w - 1;
                int dx = (r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]);
                int dy = (r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]);
                gx[(size_t)y * w + x] = (int16_t)dx; gy[(size_t)y * w + x] = (int16_t)dy;
                mag[(size_t)y * w + x] = abs(dx) + abs(dy);
            }
        }
        int sp = 0;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                size_t i = (size_t)y * w + x;
                int m = mag[i]; uint8_t s = 0;
                if (m > low) {
                    int dx = gx[i], dy = gy[i];
                    int64_t ax = abs(dx), ay = (int64_t)abs(dy) << 15;
                    int64_t t22 = ax * TG22, t67 = t22 + (ax << 16);
#define MAG(yy, xx) (((yy) < 0 || (yy) >= h || (xx) < 0 || (xx) >= w) ? 0 : mag[(size_t)(yy) * w + (xx)])
                    int keep;
                    if (ay < t22) keep = m > MAG(y, x - 1) && m >= MAG(y, x + 1);
                    else if (ay > t67) keep = m > MAG(y - 1, x) && m >= MAG(y + 1, x);
                    else { int sgn = ((dx ^ dy) < 0) ? -1 : 1; keep = m > MAG(y - 1, x - sgn) && m > MAG(y + 1, x + sgn); }
#undef MAG
                    if (keep) { s = 1; if (m > high) { s = 2; stack[sp++] = (int32_t)i; } }
                }
                st[i] = s;
            }
        while (sp > 0) {
            int32_t i = stack[--sp]; int y = i / w, x = i % w;
            for (int oy = -1; oy <= 1; ++oy) { int yy = y + oy; if (yy < 0 || yy >= h) continue;
                for (int ox = -1; ox <= 1; ++ox) { int xx = x + ox; if (xx < 0 || xx >= w) continue;
                    size_t j = (size_t)yy * w + xx;
                    if (st[j] == 1) { st[j] = 2; stack[sp++] = (int32_t)j; } } }
        }
        uint8_t* dst = out + (size_t)im * hw * (replicate3 ? 3 : 1);
        if (replicate3) for (size_t i = 0; i < hw; ++i) { uint8_t v = st[i] == 2 ? 255 : 0; dst[3*i] = v; dst[3*i+1] = v; dst[3*i+2] = v; }
        else for (size_t i = 0; i < hw; ++i) dst[i] = st[i] == 2 ? 255 : 0;
    }
    free(gray); free(mag); free(gx); free(gy); free(st); free(stack);
    return 0;
}


/* Optional Gaussian pre-stage (default OFF in the path; see include/fie_b200.h): restates cv2.GaussianBlur(img, (5, 5), 0) for
 * uint8 — OpenCV's fixed-point separable filter with the exact kernel [1 4 6 4 1]/16 per axis and a single final rounding,
 * i.e. (sum_ij w_i w_j p_ij + 128) >> 8, BORDER_REFLECT_101.  PINNED against cv2 4.13.0 (tests/test_canny_oracle.py).
 * img/out: [n][h][w][channels] u8. */
static int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
int fie_oracle_gaussian_blur5_u8(const uint8_t* img, uint8_t* out, int n, int h, int w, int channels)
{
    static const int k[5] = {1, 4, 6, 4, 1};
    if (n < 0 || h <= 0 || w <= 0 || channels <= 0) return 1;
    const size_t row = (size_t)w * channels;
    int32_t* tmp = (int32_t*)malloc((size_t)h * row * sizeof(int32_t));
    if (!tmp) return 2;
    for (int im = 0; im < n; ++im) {
        const uint8_t* src = img + (size_t)im * h * row;
        uint8_t* dst = out + (size_t)im * h * row;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                for (int c = 0; c < channels; ++c) {
                    int acc = 0;
                    for (int i = 0; i < 5; ++i) acc += k[i] * src[(size_t)y * row + (size_t)reflect101(x + i - 2, w) * channels + c];
                    tmp[(size_t)y * row + (size_t)x * channels + c] = acc;
                }
        for (int y = 0; y < h; ++y)
            for (size_t e = 0; e < row; ++e) {
                int acc = 0;
                for (int j = 0; j < 5; ++j) acc += k[j] * tmp[(size_t)reflect101(y + j - 2, h) * row + e];
                dst[(size_t)y * row + e] = (uint8_t)((acc + 128) >> 8);
            }
    }
    free(tmp);
    return 0;
}

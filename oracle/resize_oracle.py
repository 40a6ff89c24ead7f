"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of ``PIL.Image.resize(size, Image.LANCZOS)`` for 8-bit RGB images — the host-side step the reference performs at
``src/pipeline.py:251`` (``image.resize((1024, 1024), Image.LANCZOS)``) before Canny and the VAE see the pixels.  Pillow's
``ImagingResample`` (src/libImaging/Resample.c, Pillow 12.2): per output coordinate a window of double-precision Lanczos-3 weights,
normalised, converted to 2^22 fixed point with round-half-away; a horizontal pass and then a vertical pass, each accumulating in
int32 from 2^21 and clipping ``acc >> 22`` to uint8; a pass is skipped when that dimension does not change.
Pinned: bit-exact against Pillow itself on up-, down- and mixed-scale cases (tests/test_resize_oracle.py)."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x: float) -> float:
    return 1.0 if x == 0.0 else math.sin(x * math.pi) / (x * math.pi)


def _lanczos(x: float) -> float:
    return _sinc(x) * _sinc(x / 3.0) if -3.0 <= x < 3.0 else 0.0


def lanczos_coeffs(in_size: int, out_size: int):
    """-> (bounds int32 [out, 2] = (first input index, count), coefficients int32 [out, ksize], ksize)  (precompute_coeffs +
    normalize_coeffs_8bpc)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resize_lanczos_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-identical to ``Image.fromarray(img).resize((out_w, out_h), Image.LANCZOS)``."""
    h, w, _ = img.shape
    x = img.astype(np.int64)
    if out_w != w:
        b, k, _ = lanczos_coeffs(w, out_w)
        tmp = np.zeros((h, out_w, x.shape[2]), np.int64)
        for xx in range(out_w):
            x0, n = b[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(x[:, x0:x0 + n, :], k[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = tmp
    if out_h != h:
        b, k, _ = lanczos_coeffs(h, out_h)
        out = np.zeros((out_h, x.shape[1], x.shape[2]), np.int64)
        for yy in range(out_h):
            y0, n = b[yy]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(k[yy, :n].astype(np.int64), x[y0:y0 + n], axes=([0], [0]))
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        x = out
    return x.astype(np.uint8)

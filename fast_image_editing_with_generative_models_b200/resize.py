"""Host side of the GPU resamplers (``fie_resample_lanczos_u8``, ``fie_resample_f32``): the window / weight tables of Pillow's
``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (Resample.c) in double precision, cached per (input size, output size).
Replaces ``image.resize((1024, 1024), Image.LANCZOS)`` at reference ``src/pipeline.py:251`` bit for bit, so Canny and the VAE see
the pixels the reference would have fed them."""
from __future__ import annotations

import functools
import math
from typing import Tuple

import torch

PRECISION_BITS = 32 - 8 - 2


def _sinc(x: float) -> float:
    return 1.0 if x == 0.0 else math.sin(x * math.pi) / (x * math.pi)


def _lanczos(a: float) -> float:
    return _sinc(a) * _sinc(a / 3.0) if -3.0 <= a < 3.0 else 0.0


def _bicubic(x: float) -> float:
    """Pillow's bicubic_filter (Resample.c, a = -0.5): the resampler of the CLIP image processor in front of CLIPScore."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_PILLOW_FILTERS = {"lanczos": (_lanczos, 3.0), "bicubic": (_bicubic, 2.0)}


@functools.lru_cache(maxsize=64)
def pillow_tables(in_size: int, out_size: int, filter_name: str = "lanczos") -> Tuple[torch.Tensor, torch.Tensor, int]:
    """-> (bounds int32 [out, 2], coefficients int32 [out, ksize], ksize) on the CPU, for ``Image.resize`` with that filter."""
    filt, filt_support = _PILLOW_FILTERS[filter_name]
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = filt_support * fs
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / fs
    one = 1 << PRECISION_BITS
    bounds, coeffs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        x0 = max(int(center - support + 0.5), 0)
        cnt = min(int(center + support + 0.5), in_size) - x0
        w = []
        total = 0.0
        for x in range(cnt):
            v = filt((x + x0 - center + 0.5) * inv)
            w.append(v)
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        row = [int(-0.5 + v * one) if v < 0 else int(0.5 + v * one) for v in w]
        coeffs.append(row + [0] * (ksize - cnt))
        bounds.append((x0, cnt))
    return torch.tensor(bounds, dtype=torch.int32), torch.tensor(coeffs, dtype=torch.int32), ksize


def lanczos_tables(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    return pillow_tables(in_size, out_size, "lanczos")


@functools.lru_cache(maxsize=64)
def aa_bilinear_tables(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """-> (bounds int32 [out, 2], weights fp32 [out, ksize], ksize): the windows and normalised triangle weights of ATen's
    antialiased bilinear interpolation (``_upsample_bilinear2d_aa``, what ``transforms.Resize(size, antialias=True)`` runs on a float
    tensor at reference ``src/metrics.py:122,134``), computed in fp32 as ATen does for a float input."""
    import numpy as np
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = f32(1.0) * scale if scale >= 1.0 else f32(1.0)            # interp_size / 2 = 1 for the triangle filter
    ksize = int(math.ceil(float(support))) * 2 + 1
    inv = f32(1.0) / scale if scale >= 1.0 else f32(1.0)
    bounds, coeffs = [], []
    for i in range(out_size):
        center = scale * f32(i + 0.5)
        x0 = max(int(center - support + f32(0.5)), 0)
        cnt = min(int(center + support + f32(0.5)), in_size) - x0
        w = np.zeros(ksize, dtype=np.float32)
        for j in range(cnt):
            x = abs(f32(j + x0) - center + f32(0.5)) * inv
            w[j] = f32(1.0) - x if x < 1.0 else f32(0.0)
        total = w[:cnt].sum(dtype=np.float32)
        if total != 0:
            w[:cnt] /= total
        coeffs.append(w)
        bounds.append((x0, cnt))
    return torch.tensor(bounds, dtype=torch.int32), torch.from_numpy(np.stack(coeffs)), ksize

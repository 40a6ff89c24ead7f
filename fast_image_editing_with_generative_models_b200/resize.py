"""Host side of the GPU Lanczos resize (``fie_resample_lanczos_u8``): the window / weight tables of Pillow's
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


@functools.lru_cache(maxsize=64)
def lanczos_tables(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """-> (bounds int32 [out, 2], coefficients int32 [out, ksize], ksize) on the CPU."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 3.0 * fs
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
            a = (x + x0 - center + 0.5) * inv
            v = _sinc(a) * _sinc(a / 3.0) if -3.0 <= a < 3.0 else 0.0
            w.append(v)
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        row = [int(-0.5 + v * one) if v < 0 else int(0.5 + v * one) for v in w]
        coeffs.append(row + [0] * (ksize - cnt))
        bounds.append((x0, cnt))
    return torch.tensor(bounds, dtype=torch.int32), torch.tensor(coeffs, dtype=torch.int32), ksize

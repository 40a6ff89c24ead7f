"""Minimal ``MetricsCalculator`` so ``--compute_metrics`` keeps working without torchmetrics (evaluation is OUT OF SCOPE,
SURVEY 2.1 #3): SSIM / PSNR / MSE as the reference defines them (``src/metrics.py:174-176,215-239``); LPIPS / CLIP need
networks and weights that are not available offline and are reported as NaN."""
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image


def _ssim(a, b, data_range=1.0):
    k = torch.arange(11, dtype=torch.float32) - 5
    g = torch.exp(-(k ** 2) / (2 * 1.5 ** 2)); g = g / g.sum()
    w = (g[:, None] * g[None, :])[None, None].expand(3, 1, 11, 11)
    ap, bp = F.pad(a, (5,) * 4, mode="reflect"), F.pad(b, (5,) * 4, mode="reflect")
    mu_a, mu_b = F.conv2d(ap, w, groups=3), F.conv2d(bp, w, groups=3)
    s_aa = F.conv2d(ap * ap, w, groups=3) - mu_a ** 2
    s_bb = F.conv2d(bp * bp, w, groups=3) - mu_b ** 2
    s_ab = F.conv2d(ap * bp, w, groups=3) - mu_a * mu_b
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    return float((((2 * mu_a * mu_b + c1) * (2 * s_ab + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (s_aa + s_bb + c2))).mean())


class MetricsCalculator:
    def __init__(self, device="cuda"):
        self.device = device

    @staticmethod
    def _prep(img, size=512):
        t = torch.from_numpy(np.array(img.convert("RGB").resize((size, size), Image.LANCZOS))).permute(2, 0, 1).float() / 255.0
        return t[None]

    def calculate_all_metrics(self, source_img, edited_img, prompt=None):
        a, b = self._prep(source_img), self._prep(edited_img)
        mse = float(((a - b) ** 2).mean())
        return {"ssim": _ssim(a, b), "psnr": float(10 * np.log10(1.0 / max(mse, 1e-12))), "mse": mse, "lpips": float("nan"), "clip_score": float("nan")}

    def clear_memory(self):
        pass

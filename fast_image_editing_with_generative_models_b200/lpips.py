"""LPIPS with the SqueezeNet 1.1 backbone on the fie_b200 kernels — ``LearnedPerceptualImagePatchSimilarity(net_type='squeeze')`` at
reference ``src/metrics.py:180-182,239-262``.

Published algorithm (lpips 0.1 ``pretrained_networks.squeezenet`` + ``lpips.LPIPS``, which torchmetrics vendors): the input in [-1, 1]
goes through ``ScalingLayer`` ((x - shift) / scale), then ``torchvision.models.squeezenet1_1().features``; after ``features[1], [4], [7],
[9], [10], [11], [12]`` (64, 128, 256, 384, 384, 512, 512 channels) both images' activations are unit-normalised over channels
(x / (|x| + 1e-10)), squared-differenced, weighted by the non-negative 1x1 ``lin`` layer, averaged over space, and the seven layers are
summed.  Parameter names: torchvision's (``features.N.weight``, ``features.N.{squeeze, expand1x1, expand3x3}.{weight, bias}``) and lpips'
(``lin{k}.model.1.weight`` [1, C, 1, 1]).

Every convolution is a ``fie_gemm_f16`` call with bias + ReLU in its epilogue: 1x1 convolutions directly on the NHWC rows, 3x3 ones
through ``fie_im2col3x3_f16`` (the activations are 255 / 127 / 63 / 31 pixels wide, which the TMA-im2col convolution does not tile).
Squeeze outputs are zero-padded to 64 channels so that every GEMM has K a multiple of 64; the two expand branches write the two halves
of one concatenated buffer."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops, weights

Tensor = torch.Tensor

SHIFT, SCALE = (-0.030, -0.088, -0.188), (0.458, 0.448, 0.450)
FIRES = {3: (64, 16, 64, 64), 4: (128, 16, 64, 64), 6: (128, 32, 128, 128), 7: (256, 32, 128, 128), 9: (256, 48, 192, 192),
         10: (384, 48, 192, 192), 11: (384, 64, 256, 256), 12: (512, 64, 256, 256)}          # features index -> (cin, squeeze, expand1x1, expand3x3)
POOLS = (2, 5, 8)
TAPS = (1, 4, 7, 9, 10, 11, 12)
TAP_CHANNELS = (64, 128, 256, 384, 384, 512, 512)
SQ_PAD = 64


def make_lpips_params(seed: int = 51) -> Dict[str, Tensor]:
    """Seeded random-init backbone (torchvision key names, He-style scale so activations stay O(1)) + non-negative lin weights."""
    g = torch.Generator("cpu").manual_seed(seed)
    p: Dict[str, Tensor] = {}
    rn = lambda *shape, std: torch.randn(shape, generator=g) * std
    p["features.0.weight"] = rn(64, 3, 3, 3, std=(2.0 / 27) ** 0.5)
    p["features.0.bias"] = rn(64, std=0.05)
    for i, (cin, sq, e1, e3) in FIRES.items():
        p[f"features.{i}.squeeze.weight"] = rn(sq, cin, 1, 1, std=(2.0 / cin) ** 0.5)
        p[f"features.{i}.squeeze.bias"] = rn(sq, std=0.05)
        p[f"features.{i}.expand1x1.weight"] = rn(e1, sq, 1, 1, std=(2.0 / sq) ** 0.5)
        p[f"features.{i}.expand1x1.bias"] = rn(e1, std=0.05)
        p[f"features.{i}.expand3x3.weight"] = rn(e3, sq, 3, 3, std=(2.0 / (9 * sq)) ** 0.5)
        p[f"features.{i}.expand3x3.bias"] = rn(e3, std=0.05)
    for k, c in enumerate(TAP_CHANNELS):
        p[f"lin{k}.model.1.weight"] = torch.rand((1, c, 1, 1), generator=g) * (2.0 / c)
    return p


def _pad_cols(w: Tensor, k: int) -> Tensor:
    out = torch.zeros((w.shape[0], k), dtype=w.dtype)
    out[:, :w.shape[1]] = w
    return out


class LPIPSSqueeze:
    def __init__(self, params: Dict[str, Tensor], device):
        self.dev = torch.device(device)
        h16 = lambda t: t.to(self.dev, torch.float16).contiguous()
        f32 = lambda t: t.to(self.dev, torch.float32).contiguous()
        P = lambda k: params[k].float().cpu()
        # conv1: [64, 3, 3, 3] -> [64, (ky, kx, c) = 27 -> 64]
        self.w0 = h16(_pad_cols(P("features.0.weight").permute(0, 2, 3, 1).reshape(64, 27), 64))
        self.b0 = f32(P("features.0.bias"))
        self.fires: Dict[int, Tuple] = {}
        for i, (cin, sq, e1, e3) in FIRES.items():
            ws = torch.zeros((SQ_PAD, cin)); ws[:sq] = P(f"features.{i}.squeeze.weight").reshape(sq, cin)
            bs = torch.zeros((SQ_PAD,)); bs[:sq] = P(f"features.{i}.squeeze.bias")
            w1 = _pad_cols(P(f"features.{i}.expand1x1.weight").reshape(e1, sq), SQ_PAD)
            w3 = weights.pack_conv3x3(P(f"features.{i}.expand3x3.weight"), pad_cin_to=SQ_PAD)        # [e3, 9 * 64], K order (ky, kx, c)
            self.fires[i] = (h16(ws), f32(bs), h16(w1), f32(P(f"features.{i}.expand1x1.bias")), h16(w3), f32(P(f"features.{i}.expand3x3.bias")), e1, e3)
        self.lins = [f32(P(f"lin{k}.model.1.weight").reshape(-1)) for k in range(len(TAPS))]
        for k, c in enumerate(TAP_CHANNELS):
            if self.lins[k].numel() != c:
                raise ValueError(f"LPIPSSqueeze: lin{k} has {self.lins[k].numel()} weights, expected {c}")

    def _fire(self, x: Tensor, i: int) -> Tensor:
        ws, bs, w1, b1, w3, b3, e1, e3 = self.fires[i]
        n, h, w, cin = x.shape
        s = ops.gemm(x.view(-1, cin), ws, col_bias=bs, act=ops.ACT_RELU)                           # [rows, 64] (channels >= squeeze are 0)
        out = torch.empty((n, h, w, e1 + e3), dtype=torch.float16, device=x.device)
        o2 = out.view(-1, e1 + e3)
        ops.gemm(s, w1, col_bias=b1, act=ops.ACT_RELU, out=o2[:, :e1])
        ops.gemm(ops.im2col3x3(s.view(n, h, w, SQ_PAD), 1, 1), w3, col_bias=b3, act=ops.ACT_RELU, out=o2[:, e1:])
        return out

    @torch.no_grad()
    def features(self, img_u8: Tensor) -> List[Tensor]:
        """uint8 [N,H,W,3] -> the seven tapped activations (NHWC fp16)."""
        n, h, w, _ = img_u8.shape
        a = ops.im2col3x3(img_u8, 2, 0, SHIFT, SCALE, kpad=64)
        oh, ow = (h - 3) // 2 + 1, (w - 3) // 2 + 1
        x = ops.gemm(a, self.w0, col_bias=self.b0, act=ops.ACT_RELU).view(n, oh, ow, 64)
        taps = [x]
        for i in range(2, 13):
            if i in POOLS:
                x = ops.maxpool3s2_ceil(x)
            else:
                x = self._fire(x, i)
            if i in TAPS:
                taps.append(x)
        return taps

    @torch.no_grad()
    def distance(self, img0_u8: Tensor, img1_u8: Tensor) -> Tensor:
        """uint8 [N,H,W,3] pairs -> fp64 [N] LPIPS distances (the images are read as v/255*2-1, the [-1, 1] input of the reference)."""
        if img0_u8.shape != img1_u8.shape:
            raise ValueError("LPIPSSqueeze.distance: shapes differ")
        n = img0_u8.shape[0]
        taps = self.features(torch.cat([img0_u8, img1_u8], 0))
        total = torch.zeros((n,), dtype=torch.float64, device=self.dev)
        for f, lin in zip(taps, self.lins):
            total += ops.lpips_layer(f[:n].contiguous(), f[n:].contiguous(), lin)
        return total

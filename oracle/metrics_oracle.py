"""TEST INFRASTRUCTURE ONLY — CPU / torch restatement of the reference's evaluation metrics (``/root/reference/src/metrics.py:24-386``,
SURVEY 8(f)-4).  Nothing in the product imports this file.

The reference computes its metrics with third-party packages that are NOT installed here (torchmetrics >= 1.0 and its vendored lpips;
``torch.hub`` DINO needs the network), so each function restates the published algorithm and is pinned as far as the image allows:

| function | restates | pin |
|---|---|---|
| :func:`ssim` | torchmetrics ``functional/image/ssim.py::_ssim_update`` (Gaussian 11, sigma 1.5, reflect pad + crop, variance clamp) | **unpinned** against torchmetrics; checked against an independent float64 valid-window evaluation of Wang et al.'s formula and against ``scipy.ndimage.gaussian_filter(sigma=1.5, truncate=3.5)`` as the window (``tests/test_metrics_oracle.py``) |
| :func:`mse`, :func:`psnr` | ``MeanSquaredError``, ``PeakSignalNoiseRatio(data_range=1.0)`` | definitions; exact integer arithmetic cross-check |
| :func:`clip_score` | torchmetrics ``multimodal/clip_score.py::_clip_score_update`` (100 * cosine, floor 0) | towers = ``transformers.CLIPModel`` itself; the cosine equals ``CLIPModel.forward().logits_per_image / exp(logit_scale)``; image processor = Pillow bicubic (the PIL path of transformers 4.x ``CLIPImageProcessor``: shortest edge 224, centre crop, 1/255, mean / std) |
| :class:`DinoViT`, :func:`dino_distance` | facebookresearch/dino ``vision_transformer.py`` + reference ``src/metrics.py:24-148`` | pinned against ``transformers.ViTModel`` (the HF port of the same checkpoint family) with remapped weights; resize = ``torch.nn.functional.interpolate(antialias=True)`` = what torchvision ``Resize`` calls |
| :func:`lpips_squeeze` | lpips 0.1 ``LPIPS(net='squeeze')`` (what torchmetrics vendors) | backbone = ``torchvision.models.squeezenet1_1`` itself; ScalingLayer / normalize_tensor / lin / spatial_average head **unpinned** |
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
LPIPS_SHIFT, LPIPS_SCALE = (-0.030, -0.088, -0.188), (0.458, 0.448, 0.450)


def _to_float_nchw(img_u8: np.ndarray) -> torch.Tensor:
    """uint8 [H,W,3] -> float32 [1,3,H,W] / 255 (``MetricsCalculator._pil_to_tensor``, reference src/metrics.py:200-212)."""
    return torch.from_numpy(np.asarray(img_u8).astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)


# ---------------------------------------------------------------------------------------------------------------------------
def ssim(a_u8: np.ndarray, b_u8: np.ndarray, kernel_size: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03,
         dtype=torch.float32) -> float:
    """torchmetrics ``StructuralSimilarityIndexMeasure(data_range=1.0)`` on one image pair."""
    p, t = _to_float_nchw(a_u8).to(dtype), _to_float_nchw(b_u8).to(dtype)
    c = p.shape[1]
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=dtype)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    g = (g / g.sum()).unsqueeze(0)
    kernel = torch.matmul(g.t(), g).expand(c, 1, kernel_size, kernel_size)
    pad = (kernel_size - 1) // 2
    p, t = F.pad(p, (pad,) * 4, mode="reflect"), F.pad(t, (pad,) * 4, mode="reflect")
    c1, c2 = (k1 * 1.0) ** 2, (k2 * 1.0) ** 2
    out = F.conv2d(torch.cat((p, t, p * p, t * t, p * t)), kernel, groups=c)
    mu_p, mu_t, e_pp, e_tt, e_pt = out.split(1)
    mu_pp, mu_tt, mu_pt = mu_p.pow(2), mu_t.pow(2), mu_p * mu_t
    s_pp, s_tt, s_pt = torch.clamp(e_pp - mu_pp, min=0.0), torch.clamp(e_tt - mu_tt, min=0.0), e_pt - mu_pt
    full = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_pp + s_tt + c2))
    return float(full[..., pad:-pad, pad:-pad].reshape(1, -1).mean(-1))


def ssim_valid_window_f64(a_u8: np.ndarray, b_u8: np.ndarray, kernel_size: int = 11, sigma: float = 1.5) -> float:
    """Independent evaluation used to check :func:`ssim`: Wang et al. (2004) eq. 13 with Gaussian-weighted local moments over every
    fully-inside window, float64, separable filters through numpy (no padding, no torch)."""
    a, b = a_u8.astype(np.float64) / 255.0, b_u8.astype(np.float64) / 255.0
    x = np.arange(kernel_size) - (kernel_size - 1) / 2
    g = np.exp(-0.5 * (x / sigma) ** 2)
    g /= g.sum()

    def blur(m):
        h = sum(g[k] * m[:, k:m.shape[1] - kernel_size + 1 + k] for k in range(kernel_size))
        return sum(g[k] * h[k:h.shape[0] - kernel_size + 1 + k] for k in range(kernel_size))
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for ch in range(a.shape[2]):
        p, t = a[..., ch], b[..., ch]
        mp, mt = blur(p), blur(t)
        vp, vt, cov = np.maximum(blur(p * p) - mp * mp, 0), np.maximum(blur(t * t) - mt * mt, 0), blur(p * t) - mp * mt
        vals.append(((2 * mp * mt + c1) * (2 * cov + c2)) / ((mp * mp + mt * mt + c1) * (vp + vt + c2)))
    return float(np.mean(vals))


def mse(a_u8: np.ndarray, b_u8: np.ndarray) -> float:
    """torchmetrics ``MeanSquaredError`` on the flattened float images (reference src/metrics.py:310-336)."""
    d = a_u8.astype(np.int64) - b_u8.astype(np.int64)
    return float((d * d).sum()) / (255.0 * 255.0 * d.size)


def psnr(a_u8: np.ndarray, b_u8: np.ndarray) -> float:
    """torchmetrics ``PeakSignalNoiseRatio(data_range=1.0)``: 10 log10(1 / MSE)."""
    m = mse(a_u8, b_u8)
    return float("inf") if m == 0 else 10.0 * math.log10(1.0 / m)


# ---------------------------------------------------------------------------------------------------------------------------
def clip_preprocess(img_u8: np.ndarray, size: int = 224) -> torch.Tensor:
    """CLIP image processor, PIL path: bicubic resize of the shorter side to ``size``, centre crop, 1/255, normalise -> [1,3,S,S]."""
    from PIL import Image
    h, w = img_u8.shape[:2]
    oh, ow = (size, int(size * w / h)) if h <= w else (int(size * h / w), size)
    im = Image.fromarray(img_u8)
    if (oh, ow) != (h, w):
        im = im.resize((ow, oh), Image.BICUBIC)
    arr = np.asarray(im)
    top, left = (oh - size) // 2, (ow - size) // 2
    arr = arr[top:top + size, left:left + size].astype(np.float32) / 255.0
    arr = (arr - np.asarray(CLIP_MEAN, np.float32)) / np.asarray(CLIP_STD, np.float32)
    return torch.from_numpy(arr).permute(2, 0, 1).unsqueeze(0)


@torch.no_grad()
def clip_cosine(model, img_u8: np.ndarray, input_ids: torch.Tensor) -> float:
    """``model``: a ``transformers.CLIPModel``.  cos(image features, text features) = torchmetrics' score / 100 before its floor."""
    dev = next(model.parameters()).device
    px = clip_preprocess(img_u8, model.config.vision_config.image_size).to(dev)
    v = model.visual_projection(model.vision_model(pixel_values=px).pooler_output)
    t = model.text_projection(model.text_model(input_ids=input_ids.to(dev)).pooler_output)
    v, t = v / v.norm(p=2, dim=-1, keepdim=True), t / t.norm(p=2, dim=-1, keepdim=True)
    return float((v * t).sum(-1))


def clip_score(model, img_u8: np.ndarray, input_ids: torch.Tensor) -> float:
    """torchmetrics ``CLIPScore``: 100 * cosine, floored at 0."""
    return max(100.0 * clip_cosine(model, img_u8, input_ids), 0.0)


# ---------------------------------------------------------------------------------------------------------------------------
class DinoViT(torch.nn.Module):
    """facebookresearch/dino ``VisionTransformer`` (patch embed conv with bias, class token, learned position embeddings, pre-norm
    blocks with a fused qkv Linear and erf-GELU MLP, LayerNorm eps 1e-6) from a state dict with the DINO key names."""

    def __init__(self, params: Dict[str, torch.Tensor], patch_size: int, num_heads: int):
        super().__init__()
        self.p = {k: v.float() for k, v in params.items()}
        self.patch, self.heads = patch_size, num_heads
        self.depth = 1 + max(int(k.split(".")[1]) for k in params if k.startswith("blocks."))

    def to(self, dev):
        self.p = {k: v.to(dev) for k, v in self.p.items()}
        return self

    def forward(self, x: torch.Tensor, capture_qkv: bool = False):
        p = self.p
        b = x.shape[0]
        c = p["cls_token"].shape[-1]
        x = F.conv2d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=self.patch).flatten(2).transpose(1, 2)
        x = torch.cat([p["cls_token"].expand(b, -1, -1), x], 1) + p["pos_embed"]
        qkvs = []
        for i in range(self.depth):
            pre = f"blocks.{i}."
            y = F.layer_norm(x, (c,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-6)
            qkv = F.linear(y, p[pre + "attn.qkv.weight"], p[pre + "attn.qkv.bias"])
            qkvs.append(qkv)
            n = qkv.shape[1]
            q, k, v = qkv.reshape(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
            a = ((q @ k.transpose(-2, -1)) * (c // self.heads) ** -0.5).softmax(-1)
            y = (a @ v).transpose(1, 2).reshape(b, n, c)
            x = x + F.linear(y, p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"])
            y = F.layer_norm(x, (c,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-6)
            y = F.linear(F.gelu(F.linear(y, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"])), p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
            x = x + y
        x = F.layer_norm(x, (c,), p["norm.weight"], p["norm.bias"], 1e-6)
        return (x, qkvs) if capture_qkv else x


def dino_preprocess(img_u8: np.ndarray, resize_to: int = 224) -> torch.Tensor:
    """``DinoDistanceMetric._to_tensor`` (reference src/metrics.py:124-136): /255, torchvision ``Resize(resize_to, antialias=True)``
    (bilinear; the shorter side becomes ``resize_to``), ImageNet normalisation -> [1,3,h,w]."""
    x = torch.from_numpy(np.asarray(img_u8)).permute(2, 0, 1).float() / 255.0
    h, w = x.shape[1:]
    oh, ow = (resize_to, int(resize_to * w / h)) if h <= w else (int(resize_to * h / w), resize_to)
    x = F.interpolate(x.unsqueeze(0), size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)
    mean, std = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


@torch.no_grad()
def dino_keys_self_similarity(model: DinoViT, x: torch.Tensor, layer: int) -> torch.Tensor:
    """``_VitExtractor.get_keys_self_sim_from_input`` (reference src/metrics.py:72-84) -> [1, T, T]."""
    _, qkvs = model(x, capture_qkv=True)
    qkv = qkvs[layer][0]                                                      # [T, 3C] (the hook output of blocks[layer].attn.qkv)
    t, c3 = qkv.shape
    heads, dim = model.heads, c3 // 3 // model.heads
    keys = qkv.reshape(t, 3, heads, dim).permute(1, 2, 0, 3)[1]               # [heads, T, dim]
    cat = keys.transpose(0, 1).reshape(t, heads * dim)[None]
    norm = cat.norm(dim=2, keepdim=True)
    factor = torch.clamp(norm @ norm.permute(0, 2, 1), min=1e-8)
    return (cat @ cat.permute(0, 2, 1)) / factor


@torch.no_grad()
def dino_distance(model: DinoViT, src_u8: np.ndarray, edited_u8: np.ndarray, layer: int = 11, resize_to: int = 224) -> float:
    dev = next(iter(model.p.values())).device
    a = dino_keys_self_similarity(model, dino_preprocess(src_u8, resize_to).to(dev), layer)
    b = dino_keys_self_similarity(model, dino_preprocess(edited_u8, resize_to).to(dev), layer)
    return float(F.mse_loss(b, a))


# ---------------------------------------------------------------------------------------------------------------------------
LPIPS_TAPS = (1, 4, 7, 9, 10, 11, 12)


def squeezenet_features(params: Dict[str, torch.Tensor]):
    """``torchvision.models.squeezenet1_1().features`` loaded with the ``features.*`` entries of ``params``."""
    import torchvision
    net = torchvision.models.squeezenet1_1(weights=None)
    sd = net.state_dict()
    for k in sd:
        if k.startswith("features."):
            sd[k] = params[k].float()
    net.load_state_dict(sd)
    return net.features.eval()


@torch.no_grad()
def lpips_squeeze(features, lins: Sequence[torch.Tensor], img0_u8: np.ndarray, img1_u8: np.ndarray) -> float:
    """lpips 0.1 ``LPIPS(net='squeeze', lpips=True, spatial=False)`` on a [-1, 1] pair (reference src/metrics.py:239-262)."""
    dev = next(features.parameters()).device
    shift, scale = torch.tensor(LPIPS_SHIFT, device=dev).view(1, 3, 1, 1), torch.tensor(LPIPS_SCALE, device=dev).view(1, 3, 1, 1)

    def taps(img):
        x = (_to_float_nchw(img).to(dev) * 2 - 1 - shift) / scale
        out: List[torch.Tensor] = []
        for i, layer in enumerate(features):
            x = layer(x)
            if i in LPIPS_TAPS:
                out.append(x)
        return out
    total = 0.0
    for f0, f1, lin in zip(taps(img0_u8), taps(img1_u8), lins):
        n0 = f0 / (torch.sqrt(torch.sum(f0 ** 2, dim=1, keepdim=True)) + 1e-10)
        n1 = f1 / (torch.sqrt(torch.sum(f1 ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (n0 - n1) ** 2
        total += float((d * lin.to(dev).view(1, -1, 1, 1)).sum(1, keepdim=True).mean([2, 3]))
    return total

"""Seeded synthetic weights and inputs (no checkpoints, datasets or network exist in this environment).

Weights: torch default init, i.e. what diffusers' ``Model.from_config(cfg)`` would give (U(+-1/sqrt(fan_in)) for conv /
linear weights and biases), with the modules diffusers zero-initialises (ControlNet zero-convs, conditioning
``conv_out``, LoRA ``B``) given non-zero values so that every path is exercised, and two calibrated output gains
(UNet ``conv_out``, VAE decoder ``conv_out``) recorded in the configs.  Generated on the CPU generator so that the
oracle and the engine see bit-identical fp32 master weights on any machine.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from .configs import ControlNetConfig, UNetConfig, VAEConfig, skip_channels

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# ----------------------------------------------------------------------------------------------
# Synthetic weights (torch default init as Model.from_config would give)
# ----------------------------------------------------------------------------------------------


class _Init:
    """Collects (name, shape, bound, offset) specs, then materialises them in parallel: tensor i of a model with seed
    ``s`` is ``offset + U(-bound, bound)`` drawn from its own CPU generator seeded ``s * 1000003 + i`` (order-independent,
    so generation can use every host core)."""
    meta = False  # class-level switch: shape-only parameters on the meta device (for counting)

    def __init__(self, seed: int):
        self.seed = seed
        self.specs = []
        self.p: Params = {}

    def _add(self, name, shape, bound, offset=0.0):
        self.specs.append((name, tuple(shape), float(bound), float(offset)))
        self.p[name] = None

    def uniform(self, shape, bound):
        """Immediate draw (used by the LoRA recipe)."""
        if _Init.meta:
            return torch.empty(shape, device="meta")
        g = torch.Generator("cpu").manual_seed(self.seed * 1000003 + len(self.specs) + 500000)
        self.specs.append(("<anon>", tuple(shape), bound, 0.0))
        return torch.empty(shape, dtype=torch.float32).uniform_(-bound, bound, generator=g)

    def conv(self, name, cin, cout, k):
        b = 1.0 / math.sqrt(cin * k * k)
        self._add(name + ".weight", (cout, cin, k, k), b)
        self._add(name + ".bias", (cout,), b)

    def linear(self, name, cin, cout, bias=True):
        b = 1.0 / math.sqrt(cin)
        self._add(name + ".weight", (cout, cin), b)
        if bias:
            self._add(name + ".bias", (cout,), b)

    def norm(self, name, c):
        # gamma=1, beta=0 is the from_config default; a seeded perturbation exercises the affine path.
        self._add(name + ".weight", (c,), 0.1, 1.0)
        self._add(name + ".bias", (c,), 0.1)

    def finish(self) -> Params:
        if _Init.meta:
            for name, shape, _, _ in self.specs:
                if name != "<anon>":
                    self.p[name] = torch.empty(shape, device="meta")
            return self.p

        def make(item):
            i, (name, shape, bound, offset) = item
            g = torch.Generator("cpu").manual_seed(self.seed * 1000003 + i)
            t = torch.empty(shape, dtype=torch.float32).uniform_(-bound, bound, generator=g)
            if offset:
                t += offset
            return name, t

        import os
        from concurrent.futures import ThreadPoolExecutor
        items = [(i, sp) for i, sp in enumerate(self.specs) if sp[0] != "<anon>"]
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
            for name, t in ex.map(make, items):
                self.p[name] = t
        return self.p


def _init_resnet(I: _Init, pre, cin, cout, temb_dim):
    I.norm(pre + ".norm1", cin)
    I.conv(pre + ".conv1", cin, cout, 3)
    if temb_dim:
        I.linear(pre + ".time_emb_proj", temb_dim, cout)
    I.norm(pre + ".norm2", cout)
    I.conv(pre + ".conv2", cout, cout, 3)
    if cin != cout:
        I.conv(pre + ".conv_shortcut", cin, cout, 1)


def _init_transformer(I: _Init, pre, c, depth, ctx_dim):
    I.norm(pre + ".norm", c)
    I.linear(pre + ".proj_in", c, c)
    for k in range(depth):
        b = f"{pre}.transformer_blocks.{k}"
        I.norm(b + ".norm1", c)
        for nm in ("to_q", "to_k", "to_v"):
            I.linear(f"{b}.attn1.{nm}", c, c, bias=False)
        I.linear(b + ".attn1.to_out.0", c, c)
        I.norm(b + ".norm2", c)
        I.linear(b + ".attn2.to_q", c, c, bias=False)
        I.linear(b + ".attn2.to_k", ctx_dim, c, bias=False)
        I.linear(b + ".attn2.to_v", ctx_dim, c, bias=False)
        I.linear(b + ".attn2.to_out.0", c, c)
        I.norm(b + ".norm3", c)
        I.linear(b + ".ff.net.0.proj", c, 8 * c)
        I.linear(b + ".ff.net.2", 4 * c, c)
    I.linear(pre + ".proj_out", c, c)


def _init_encoder_part(I: _Init, cfg: UNetConfig):
    ch = cfg.block_out_channels
    T = cfg.time_embed_dim
    I.conv("conv_in", cfg.in_channels, ch[0], 3)
    I.linear("time_embedding.linear_1", ch[0], T)
    I.linear("time_embedding.linear_2", T, T)
    I.linear("add_embedding.linear_1", cfg.projection_class_embeddings_input_dim, T)
    I.linear("add_embedding.linear_2", T, T)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _init_resnet(I, f"down_blocks.{i}.resnets.{j}", cin, cout, T)
            cin = cout
            if len(cfg.down_depths[i]):
                _init_transformer(I, f"down_blocks.{i}.attentions.{j}", cout, cfg.down_depths[i][j], cfg.cross_attention_dim)
        if i < len(ch) - 1:
            I.conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    c = ch[-1]
    _init_resnet(I, "mid_block.resnets.0", c, c, T)
    if cfg.mid_depth is not None:
        _init_transformer(I, "mid_block.attentions.0", c, cfg.mid_depth, cfg.cross_attention_dim)
        _init_resnet(I, "mid_block.resnets.1", c, c, T)


def make_unet_params(cfg: UNetConfig) -> Params:
    I = _Init(cfg.seed)
    _init_encoder_part(I, cfg)
    ch = cfg.block_out_channels
    T = cfg.time_embed_dim
    skips = skip_channels(cfg)
    rev = list(reversed(ch))
    prev = ch[-1]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            sc = skips.pop()
            _init_resnet(I, f"up_blocks.{i}.resnets.{j}", prev + sc, cout, T)
            prev = cout
            if len(cfg.up_depths[i]):
                _init_transformer(I, f"up_blocks.{i}.attentions.{j}", cout, cfg.up_depths[i][j], cfg.cross_attention_dim)
        if i < len(rev) - 1:
            I.conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    I.norm("conv_norm_out", ch[0])
    I.conv("conv_out", ch[0], cfg.out_channels, 3)
    I.finish()
    I.p["conv_out.weight"] *= cfg.conv_out_gain
    I.p["conv_out.bias"] *= cfg.conv_out_gain
    return I.p


def make_controlnet_params(cfg: ControlNetConfig) -> Params:
    I = _Init(cfg.seed)
    _init_encoder_part(I, cfg.unet)
    cc = cfg.cond_channels
    I.conv("controlnet_cond_embedding.conv_in", 3, cc[0], 3)
    for i in range(len(cc) - 1):
        I.conv(f"controlnet_cond_embedding.blocks.{2 * i}", cc[i], cc[i], 3)
        I.conv(f"controlnet_cond_embedding.blocks.{2 * i + 1}", cc[i], cc[i + 1], 3)
    # diffusers zero-inits conv_out and the zero-convs; the synthetic recipe uses default init so the
    # residual path is non-zero.
    I.conv("controlnet_cond_embedding.conv_out", cc[-1], cfg.unet.block_out_channels[0], 3)
    for i, c in enumerate(skip_channels(cfg.unet)):
        I.conv(f"controlnet_down_blocks.{i}", c, c, 1)
    c = cfg.unet.block_out_channels[-1]
    I.conv("controlnet_mid_block", c, c, 1)
    return I.finish()


def _init_vae_resnet(I, pre, cin, cout):
    _init_resnet(I, pre, cin, cout, 0)


def _init_vae_attn(I, pre, c):
    I.norm(pre + ".group_norm", c)
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        I.linear(f"{pre}.{nm}", c, c)


def make_vae_params(cfg: VAEConfig) -> Params:
    I = _Init(cfg.seed)
    ch = cfg.block_out_channels
    L = cfg.latent_channels
    I.conv("encoder.conv_in", 3, ch[0], 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _init_vae_resnet(I, f"encoder.down_blocks.{i}.resnets.{j}", cin, cout)
            cin = cout
        if i < len(ch) - 1:
            I.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    c = ch[-1]
    _init_vae_resnet(I, "encoder.mid_block.resnets.0", c, c)
    _init_vae_attn(I, "encoder.mid_block.attentions.0", c)
    _init_vae_resnet(I, "encoder.mid_block.resnets.1", c, c)
    I.norm("encoder.conv_norm_out", c)
    I.conv("encoder.conv_out", c, 2 * L, 3)
    I.conv("quant_conv", 2 * L, 2 * L, 1)
    I.conv("post_quant_conv", L, L, 1)
    I.conv("decoder.conv_in", L, c, 3)
    _init_vae_resnet(I, "decoder.mid_block.resnets.0", c, c)
    _init_vae_attn(I, "decoder.mid_block.attentions.0", c)
    _init_vae_resnet(I, "decoder.mid_block.resnets.1", c, c)
    rev = list(reversed(ch))
    cin = c
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            _init_vae_resnet(I, f"decoder.up_blocks.{i}.resnets.{j}", cin, cout)
            cin = cout
        if i < len(rev) - 1:
            I.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    I.norm("decoder.conv_norm_out", ch[0])
    I.conv("decoder.conv_out", ch[0], 3, 3)
    I.finish()
    I.p["decoder.conv_out.weight"] *= cfg.conv_out_gain
    I.p["decoder.conv_out.bias"] *= cfg.conv_out_gain
    return I.p


# ---- LCM-LoRA (peft) -------------------------------------------------------------------------

LORA_TARGET_SUFFIXES = ("to_q", "to_k", "to_v", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2",
                        "conv1", "conv2", "conv_shortcut", "downsamplers.0.conv", "upsamplers.0.conv", "time_emb_proj")


def make_lora_params(unet_params: Params, rank: int = 64, seed: int = 16, b_scale: float = 0.02) -> Params:
    """Synthetic LCM-LoRA (r=64, alpha=64): A default-init, B small non-zero (peft zero-inits B)."""
    I = _Init(seed)
    out: Params = {}
    for k, w in unet_params.items():
        if not k.endswith(".weight"):
            continue
        base = k[: -len(".weight")]
        if not base.endswith(LORA_TARGET_SUFFIXES) or w.dim() < 2:
            continue
        cout, cin = w.shape[0], w.shape[1]
        if w.dim() == 4:
            kk = w.shape[2]
            out[base + ".lora_A.weight"] = I.uniform((rank, cin, kk, kk), 1.0 / math.sqrt(cin * kk * kk))
            out[base + ".lora_B.weight"] = I.uniform((cout, rank, 1, 1), b_scale / math.sqrt(rank))
        else:
            out[base + ".lora_A.weight"] = I.uniform((rank, cin), 1.0 / math.sqrt(cin))
            out[base + ".lora_B.weight"] = I.uniform((cout, rank), b_scale / math.sqrt(rank))
    return out


class shapes_only:
    """Context manager: make_*_params() return meta tensors (no memory) — for parameter counting."""

    def __enter__(self):
        _Init.meta = True

    def __exit__(self, *a):
        _Init.meta = False


def count_params(p: Params) -> int:
    return sum(v.numel() for v in p.values())


def to_dtype(p: Params, dtype, device=None) -> Params:
    return {k: v.to(device=device, dtype=dtype) for k, v in p.items()}




# ----------------------------------------------------------------------------------------------
# Synthetic inputs
# ----------------------------------------------------------------------------------------------


def synthetic_image(seed: int, h: int = 1024, w: int = 1024, kind: str = "shapes") -> np.ndarray:
    """Seeded synthetic uint8 RGB test images: image-like 'shapes' or iid 'noise'."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if kind == "smooth":
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.stack([127 + 120 * np.sin(xx / (2.0 + seed) + yy / 3.0),
                        127 + 120 * np.cos(xx / 2.5 - yy / (1.5 + seed)),
                        127 + 100 * np.sin((xx + yy) / 3.1)], axis=2)
        return np.clip(img, 0, 255).astype(np.uint8)
    img = np.full((h, w, 3), 128.0, np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(60):
        col = rng.integers(0, 256, size=3).astype(np.float32)
        if rng.random() < 0.5:
            cy, cx = rng.integers(0, h), rng.integers(0, w)
            r = int(rng.integers(max(2, h // 64), max(3, h // 6)))
            mask = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        else:
            y0, x0 = rng.integers(0, h), rng.integers(0, w)
            hh, ww = rng.integers(max(2, h // 64), max(3, h // 4)), rng.integers(max(2, w // 64), max(3, w // 4))
            mask = (yy >= y0) & (yy < y0 + hh) & (xx >= x0) & (xx < x0 + ww)
        img[mask] = col
    # separable 5-tap binomial blur (~sigma 1.0) + noise, no scipy dependency
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    p = np.pad(img, ((2, 2), (0, 0), (0, 0)), mode="edge")
    img = sum(k[i] * p[i:i + h] for i in range(5))
    p = np.pad(img, ((0, 0), (2, 2), (0, 0)), mode="edge")
    img = sum(k[i] * p[:, i:i + w] for i in range(5))
    img = img + rng.normal(0, 6, size=img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_prompt(index: int, ctx_dim: int = 2048, pooled_dim: int = 1280, tokens: int = 77):
    """Stand-in for the two CLIP text encoders (out of scope): prompt_embeds [2,77,D] and pooled
    [2,P] ~ N(0,1) from a CPU generator seeded 1000+index (row 0 = negative prompt, row 1 = positive), fp16."""
    g = torch.Generator("cpu").manual_seed(1000 + index)
    pe = torch.randn((2, tokens, ctx_dim), generator=g).to(torch.float16)
    pl = torch.randn((2, pooled_dim), generator=g).to(torch.float16)
    return pe, pl


def synthetic_noises(index: int, batch: int, h: int, w: int, count: int = 4):
    """The Gaussian draws of one edit in the reference generator's order (posterior sample xi, init noise n, then
    one per non-final step), each [batch,4,h,w] fp32 rounded to fp16 values, CPU generator seeded 2000+index."""
    g = torch.Generator("cpu").manual_seed(2000 + index)
    return [torch.randn((batch, 4, h, w), generator=g).to(torch.float16).float() for _ in range(count)]

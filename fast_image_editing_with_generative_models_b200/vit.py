"""Vision transformers of the evaluation metrics on the fie_b200 kernels — SURVEY 8(f)-4 (reference ``src/metrics.py``):

* the CLIP ViT-B/16 image tower behind ``CLIPScore("openai/clip-vit-base-patch16")`` (``src/metrics.py:185-187,264-283``), parameter
  names of transformers' ``CLIPModel`` (``vision_model.*``, ``visual_projection.weight``), and
* DINO ViT-B/8 (``torch.hub.load('facebookresearch/dino:main', 'dino_vitb8')``, ``src/metrics.py:24-33``), parameter names of the DINO
  ``VisionTransformer`` (``cls_token, pos_embed, patch_embed.proj, blocks.N.{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}, norm``),
  of which the metric needs the *keys* of one block (the ``attn.qkv`` forward hook at ``src/metrics.py:36-76``).

Patch embedding = ``fie_patchify_f16`` + one GEMM; per block LayerNorm -> fused q/k/v GEMM -> flash attention (head_dim 64, no mask)
-> out-proj GEMM (+residual) -> LayerNorm -> fc1 GEMM (activation in the epilogue) -> fc2 GEMM (+residual): the same calls as
``text_encoder.py``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops

Tensor = torch.Tensor

CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@dataclass
class ViTConfig:
    name: str = "clip-vit-b16-vision"
    style: str = "clip"                    # "clip" (transformers CLIPVisionModel keys) or "dino" (DINO VisionTransformer keys)
    image_size: int = 224
    patch_size: int = 16
    hidden_size: int = 768
    num_layers: int = 12
    num_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: Optional[int] = 512    # CLIP visual_projection; None for DINO
    seed: int = 41

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2


def clip_b16_vision_config() -> ViTConfig:
    return ViTConfig()


def dino_vitb8_config() -> ViTConfig:
    return ViTConfig(name="dino-vitb8", style="dino", patch_size=8, hidden_act="gelu", layer_norm_eps=1e-6, projection_dim=None, seed=42)


def tiny_vit_config(style: str = "clip", image_size: int = 64, patch_size: int = 16) -> ViTConfig:
    return ViTConfig(name=f"vit-tiny-{style}", style=style, image_size=image_size, patch_size=patch_size, hidden_size=128, num_layers=3, num_heads=2,
                     intermediate_size=256, hidden_act="quick_gelu" if style == "clip" else "gelu", layer_norm_eps=1e-5 if style == "clip" else 1e-6,
                     projection_dim=64 if style == "clip" else None, seed=43)


def dino_config_from_params(params: Dict[str, Tensor], name: str = "dino-vit") -> ViTConfig:
    """The architecture of a DINO ``VisionTransformer`` checkpoint read off its tensor shapes (dino_vits8 / vits16 / vitb8 / vitb16 all
    have 64-wide heads)."""
    c = params["cls_token"].shape[-1]
    ps = params["patch_embed.proj.weight"].shape[-1]
    side = int(round((params["pos_embed"].shape[1] - 1) ** 0.5))
    layers = 1 + max(int(k.split(".")[1]) for k in params if k.startswith("blocks."))
    return ViTConfig(name=name, style="dino", image_size=side * ps, patch_size=ps, hidden_size=c, num_layers=layers, num_heads=c // 64,
                     intermediate_size=params["blocks.0.mlp.fc1.weight"].shape[0], hidden_act="gelu", layer_norm_eps=1e-6, projection_dim=None)


def make_vit_params(cfg: ViTConfig) -> Dict[str, Tensor]:
    """Seeded random-init state dict with the key names of the style's published checkpoint (fp32, CPU)."""
    g = torch.Generator("cpu").manual_seed(cfg.seed)
    c, ps, p = cfg.hidden_size, cfg.patch_size, {}
    rn = lambda *shape, std=0.02: torch.randn(shape, generator=g) * std
    t = cfg.num_patches + 1
    if cfg.style == "clip":
        p["vision_model.embeddings.class_embedding"] = rn(c, std=0.05)
        p["vision_model.embeddings.patch_embedding.weight"] = rn(c, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
        p["vision_model.embeddings.position_embedding.weight"] = rn(t, c, std=0.02)
        for n in ("pre_layrnorm", "post_layernorm"):
            p[f"vision_model.{n}.weight"] = 1.0 + rn(c, std=0.05)
            p[f"vision_model.{n}.bias"] = rn(c, std=0.05)
        for i in range(cfg.num_layers):
            pre = f"vision_model.encoder.layers.{i}."
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                p[pre + f"self_attn.{n}.weight"] = rn(c, c, std=c ** -0.5)
                p[pre + f"self_attn.{n}.bias"] = rn(c, std=0.02)
            for n in ("layer_norm1", "layer_norm2"):
                p[pre + n + ".weight"] = 1.0 + rn(c, std=0.05)
                p[pre + n + ".bias"] = rn(c, std=0.05)
            p[pre + "mlp.fc1.weight"] = rn(cfg.intermediate_size, c, std=c ** -0.5)
            p[pre + "mlp.fc1.bias"] = rn(cfg.intermediate_size, std=0.02)
            p[pre + "mlp.fc2.weight"] = rn(c, cfg.intermediate_size, std=cfg.intermediate_size ** -0.5)
            p[pre + "mlp.fc2.bias"] = rn(c, std=0.02)
        if cfg.projection_dim:
            p["visual_projection.weight"] = rn(cfg.projection_dim, c, std=c ** -0.5)
    else:
        p["cls_token"] = rn(1, 1, c, std=0.05)
        p["pos_embed"] = rn(1, t, c, std=0.02)
        p["patch_embed.proj.weight"] = rn(c, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
        p["patch_embed.proj.bias"] = rn(c, std=0.02)
        for i in range(cfg.num_layers):
            pre = f"blocks.{i}."
            for n in ("norm1", "norm2"):
                p[pre + n + ".weight"] = 1.0 + rn(c, std=0.05)
                p[pre + n + ".bias"] = rn(c, std=0.05)
            p[pre + "attn.qkv.weight"] = rn(3 * c, c, std=c ** -0.5)
            p[pre + "attn.qkv.bias"] = rn(3 * c, std=0.02)
            p[pre + "attn.proj.weight"] = rn(c, c, std=c ** -0.5)
            p[pre + "attn.proj.bias"] = rn(c, std=0.02)
            p[pre + "mlp.fc1.weight"] = rn(cfg.intermediate_size, c, std=c ** -0.5)
            p[pre + "mlp.fc1.bias"] = rn(cfg.intermediate_size, std=0.02)
            p[pre + "mlp.fc2.weight"] = rn(c, cfg.intermediate_size, std=cfg.intermediate_size ** -0.5)
            p[pre + "mlp.fc2.bias"] = rn(c, std=0.02)
        p["norm.weight"] = 1.0 + rn(c, std=0.05)
        p["norm.bias"] = rn(c, std=0.05)
    return p


class VisionTransformer:
    """One ViT on one GPU.  ``embed(images)`` -> projected class-token features (CLIP ``get_image_features``); ``keys(images, layer)``
    -> the key projections of every token in block ``layer`` [N, T, C] (DINO)."""

    def __init__(self, params: Dict[str, Tensor], cfg: ViTConfig, device):
        if cfg.hidden_size // cfg.num_heads != 64 or cfg.hidden_size % 64:
            raise ValueError("VisionTransformer: head_dim must be 64")
        self.cfg, self.dev = cfg, torch.device(device)
        c = cfg.hidden_size
        h16 = lambda t: t.to(self.dev, torch.float16).contiguous()
        f32 = lambda t: t.to(self.dev, torch.float32).contiguous()
        P = lambda k: params[k]
        clip = cfg.style == "clip"
        # conv [C, 3, P, P] -> GEMM weight [C, (py, px, c)], the K order of fie_patchify_f16
        wp = P("vision_model.embeddings.patch_embedding.weight" if clip else "patch_embed.proj.weight")
        self.w_patch = h16(wp.permute(0, 2, 3, 1).reshape(c, -1))
        self.b_patch = None if clip else f32(P("patch_embed.proj.bias"))
        self.cls = h16(P("vision_model.embeddings.class_embedding") if clip else P("cls_token").reshape(c))
        self.pos = h16(P("vision_model.embeddings.position_embedding.weight") if clip else P("pos_embed").reshape(-1, c))
        if self.pos.shape[0] != cfg.num_patches + 1:
            raise ValueError(f"VisionTransformer: {self.pos.shape[0]} position embeddings for {cfg.num_patches} patches + class token "
                             "(position-embedding interpolation is not implemented: feed images of the configured size)")
        self.pre_ln = (f32(P("vision_model.pre_layrnorm.weight")), f32(P("vision_model.pre_layrnorm.bias"))) if clip else None
        self.layers: List[dict] = []
        for i in range(cfg.num_layers):
            if clip:
                pre = f"vision_model.encoder.layers.{i}."
                wqkv = torch.cat([P(pre + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")], 0)
                bqkv = torch.cat([P(pre + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")], 0)
                names = dict(ln1="layer_norm1", ln2="layer_norm2", wo="self_attn.out_proj", fc1="mlp.fc1", fc2="mlp.fc2")
            else:
                pre = f"blocks.{i}."
                wqkv, bqkv = P(pre + "attn.qkv.weight"), P(pre + "attn.qkv.bias")
                names = dict(ln1="norm1", ln2="norm2", wo="attn.proj", fc1="mlp.fc1", fc2="mlp.fc2")
            g = lambda n, s: P(pre + names[n] + "." + s)
            self.layers.append(dict(ln1=(f32(g("ln1", "weight")), f32(g("ln1", "bias"))), ln2=(f32(g("ln2", "weight")), f32(g("ln2", "bias"))),
                                    wqkv=h16(wqkv), bqkv=f32(bqkv), wo=h16(g("wo", "weight")), bo=f32(g("wo", "bias")),
                                    w1=h16(g("fc1", "weight")), b1=f32(g("fc1", "bias")), w2=h16(g("fc2", "weight")), b2=f32(g("fc2", "bias"))))
        self.post_ln = (f32(P("vision_model.post_layernorm.weight")), f32(P("vision_model.post_layernorm.bias"))) if clip else (f32(P("norm.weight")), f32(P("norm.bias")))
        self.proj = h16(P("visual_projection.weight")) if (clip and cfg.projection_dim) else None
        self.act = ops.ACT_QUICKGELU if cfg.hidden_act == "quick_gelu" else ops.ACT_GELU

    def _tokens(self, images: Tensor, mean, std) -> Tensor:
        cfg = self.cfg
        n, h, w, _ = images.shape
        if h != cfg.image_size or w != cfg.image_size:
            raise ValueError(f"VisionTransformer: images must be {cfg.image_size} x {cfg.image_size} (got {h} x {w})")
        rows = ops.patchify(images, cfg.patch_size, mean, std)
        x = ops.gemm(rows, self.w_patch, col_bias=self.b_patch)
        tok = ops.vit_assemble(x, self.cls, self.pos, n)
        if self.pre_ln is not None:
            tok = ops.layernorm(tok, *self.pre_ln, eps=cfg.layer_norm_eps)
        return tok

    def _block(self, h: Tensor, L: dict, n: int, t: int, keys_only: bool = False) -> Tensor:
        cfg = self.cfg
        c = cfg.hidden_size
        n1 = ops.layernorm(h, *L["ln1"], eps=cfg.layer_norm_eps)
        qkv = ops.gemm(n1, L["wqkv"], col_bias=L["bqkv"])
        if keys_only:
            return qkv[:, c:2 * c]
        a = ops.attention_d64(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], n, cfg.num_heads, t, t)
        h = ops.gemm(a, L["wo"], col_bias=L["bo"], residual=h)
        n2 = ops.layernorm(h, *L["ln2"], eps=cfg.layer_norm_eps)
        m = ops.gemm(n2, L["w1"], col_bias=L["b1"], act=self.act)
        return ops.gemm(m, L["w2"], col_bias=L["b2"], residual=h)

    @torch.no_grad()
    def embed(self, images: Tensor, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None) -> Tensor:
        """images uint8 (normalised here with mean / std) or already-normalised fp32 [N,S,S,3] -> fp16 [N, projection_dim or C]:
        LayerNorm of the class token after the last block (``pooler_output``), then ``visual_projection`` when there is one."""
        n, t = images.shape[0], self.cfg.num_patches + 1
        h = self._tokens(images, mean, std)
        for L in self.layers:
            h = self._block(h, L, n, t)
        cls_rows = h.view(n, t, -1)[:, 0].contiguous()
        pooled = ops.layernorm(cls_rows, *self.post_ln, eps=self.cfg.layer_norm_eps)
        return ops.gemm(pooled, self.proj) if self.proj is not None else pooled

    @torch.no_grad()
    def keys(self, images: Tensor, layer: int, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None) -> Tensor:
        """-> fp16 view [N*T, C] (row stride 3C): the key third of block ``layer``'s fused q/k/v projection, heads concatenated
        (``keys.transpose(0, 1).reshape(tokens, heads * dim)`` at reference ``src/metrics.py:72-76``)."""
        if not 0 <= layer < self.cfg.num_layers:
            raise ValueError(f"VisionTransformer.keys: layer {layer} out of range")
        n, t = images.shape[0], self.cfg.num_patches + 1
        h = self._tokens(images, mean, std)
        for L in self.layers[:layer]:
            h = self._block(h, L, n, t)
        return self._block(h, self.layers[layer], n, t, keys_only=True)

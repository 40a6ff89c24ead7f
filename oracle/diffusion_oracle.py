"""TEST INFRASTRUCTURE — plain-PyTorch restatement of the diffusers modules on the reference hot path.

The reference (``src/pipeline.py``) executes every FLOP of ``FastEditor.edit`` inside the third-party
package ``diffusers`` (unpinned ``>=0.25.0`` in the reference's ``requirements.txt:6``; the authors ran
0.35.2, reference ``README.md:350``), which is NOT vendored under ``/root/reference`` and is not installable
in this image (no wheel, no network).  This file restates the published algorithm of the modules the
reference instantiates, anchored on the reference's own call sites:

* ``StableDiffusionXLControlNetImg2ImgPipeline.__call__`` — called at ``src/pipeline.py:261-272``
  -> :func:`edit_pipeline`
* ``UNet2DConditionModel`` (SDXL at ``src/pipeline.py:147-153``, SSD-1B LCM UNet at ``:115-124``)
  -> :func:`unet_forward`
* ``ControlNetModel`` (``src/pipeline.py:82-92``) -> :func:`controlnet_forward`
* ``AutoencoderKL`` (``src/pipeline.py:94-105``) -> :func:`vae_encode_moments`, :func:`vae_decode`
* ``LCMScheduler`` (``src/pipeline.py:138-141,158-161``) -> :class:`LCMSchedule`
* LCM-LoRA (``src/pipeline.py:154``, unfused at runtime in the reference) -> ``_LoRA`` (applied inside ``_linear`` / ``_conv``)

Only ``torch.nn.functional`` ops are used (conv2d, linear, group_norm, layer_norm, silu, gelu,
scaled_dot_product_attention, interpolate(nearest), pad, cat) — the same ATen ops diffusers issues.
Parameter dictionaries use the diffusers state-dict key names so a real checkpoint can be loaded.

Parity status: PARITY UNPINNED against diffusers itself (the reference has no tests, golden vectors or
fixtures for this path and diffusers cannot be imported here).  What IS pinned offline: parameter totals
(SDXL UNet 2567.5 M, SSD-1B 1331.3 M, ControlNet-full 1251.0 M, VAE 83.7 M), the LCM timestep table
[999, 759, 499, 259] and the scheduler constants (tests/test_oracle_diffusion.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# Configs and the seeded synthetic-weight recipe are shared with the product package (data only, no kernels).
from fast_image_editing_with_generative_models_b200.configs import (  # noqa: F401
    ControlNetConfig, UNetConfig, VAEConfig, controlnet_config, sdxl_unet_config, skip_channels, ssd1b_unet_config,
    tiny_controlnet_config, tiny_unet_config, tiny_vae_config)
from fast_image_editing_with_generative_models_b200.synthetic import (  # noqa: F401
    count_params, make_controlnet_params, make_lora_params, make_unet_params, make_vae_params, shapes_only, to_dtype)

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# Modules (functional)
# ----------------------------------------------------------------------------------------------


class _LoRA:
    """Holds optional unfused LoRA weights: y = W x + (alpha/r) * B(A x)   (reference src/pipeline.py:154)."""

    def __init__(self, lora: Optional[Params] = None, scale: float = 1.0):
        self.lora = lora
        self.scale = scale


def _linear(p: Params, name: str, x: Tensor, lora: Optional[_LoRA] = None) -> Tensor:
    y = F.linear(x, p[name + ".weight"], p.get(name + ".bias"))
    if lora is not None and lora.lora is not None and name + ".lora_A.weight" in lora.lora:
        y = y + lora.scale * F.linear(F.linear(x, lora.lora[name + ".lora_A.weight"]), lora.lora[name + ".lora_B.weight"])
    return y


def _conv(p: Params, name: str, x: Tensor, stride=1, padding=1, lora: Optional[_LoRA] = None) -> Tensor:
    y = F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=stride, padding=padding)
    if lora is not None and lora.lora is not None and name + ".lora_A.weight" in lora.lora:
        a = F.conv2d(x, lora.lora[name + ".lora_A.weight"], None, stride=stride, padding=padding)
        y = y + lora.scale * F.conv2d(a, lora.lora[name + ".lora_B.weight"])
    return y


def _gn(p: Params, name: str, x: Tensor, groups: int, eps: float) -> Tensor:
    return F.group_norm(x, groups, p[name + ".weight"], p[name + ".bias"], eps)


def _ln(p: Params, name: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-5)


def resnet_block(p: Params, pre: str, x: Tensor, temb: Optional[Tensor], groups: int, eps: float,
                 lora: Optional[_LoRA] = None) -> Tensor:
    """diffusers ResnetBlock2D (output_scale_factor 1, dropout 0)."""
    h = F.silu(_gn(p, pre + ".norm1", x, groups, eps))
    h = _conv(p, pre + ".conv1", h, lora=lora)
    if temb is not None and pre + ".time_emb_proj.weight" in p:
        h = h + _linear(p, pre + ".time_emb_proj", F.silu(temb), lora)[:, :, None, None]
    h = F.silu(_gn(p, pre + ".norm2", h, groups, eps))
    h = _conv(p, pre + ".conv2", h, lora=lora)
    if pre + ".conv_shortcut.weight" in p:
        x = _conv(p, pre + ".conv_shortcut", x, padding=0, lora=lora)
    return x + h


def attention(p: Params, pre: str, x: Tensor, ctx: Optional[Tensor], head_dim: int, lora=None) -> Tensor:
    """diffusers Attention + AttnProcessor2_0 (no mask, scale 1/sqrt(d))."""
    ctx = x if ctx is None else ctx
    q = _linear(p, pre + ".to_q", x, lora)
    k = _linear(p, pre + ".to_k", ctx, lora)
    v = _linear(p, pre + ".to_v", ctx, lora)
    B, N, C = q.shape
    h = C // head_dim
    q = q.view(B, N, h, head_dim).transpose(1, 2)
    k = k.view(B, -1, h, head_dim).transpose(1, 2)
    v = v.view(B, -1, h, head_dim).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, N, C)
    return _linear(p, pre + ".to_out.0", o, lora)


def transformer_block(p: Params, pre: str, x: Tensor, ctx: Tensor, head_dim: int, lora=None) -> Tensor:
    x = x + attention(p, pre + ".attn1", _ln(p, pre + ".norm1", x), None, head_dim, lora)
    x = x + attention(p, pre + ".attn2", _ln(p, pre + ".norm2", x), ctx, head_dim, lora)
    h = _linear(p, pre + ".ff.net.0.proj", _ln(p, pre + ".norm3", x), lora)
    val, gate = h.chunk(2, dim=-1)
    h = val * F.gelu(gate)
    return x + _linear(p, pre + ".ff.net.2", h, lora)


def transformer_2d(p: Params, pre: str, x: Tensor, ctx: Tensor, depth: int, groups: int, head_dim: int, lora=None) -> Tensor:
    B, C, H, W = x.shape
    res = x
    h = _gn(p, pre + ".norm", x, groups, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _linear(p, pre + ".proj_in", h, lora)
    for k in range(depth):
        h = transformer_block(p, f"{pre}.transformer_blocks.{k}", h, ctx, head_dim, lora)
    h = _linear(p, pre + ".proj_out", h, lora)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return h + res


def sincos_embedding(t: Tensor, dim: int) -> Tensor:
    """diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): cat([cos, sin])."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_and_add_embedding(p: Params, cfg: UNetConfig, t: Tensor, text_embeds: Tensor, time_ids: Tensor, dtype) -> Tensor:
    B = text_embeds.shape[0]
    t = t.reshape(-1).expand(B)
    t_emb = sincos_embedding(t, cfg.block_out_channels[0]).to(dtype)
    emb = _linear(p, "time_embedding.linear_2", F.silu(_linear(p, "time_embedding.linear_1", t_emb)))
    tid = sincos_embedding(time_ids.reshape(-1), cfg.addition_time_embed_dim).reshape(B, -1)
    add = torch.cat([text_embeds, tid.to(dtype)], dim=-1)
    aug = _linear(p, "add_embedding.linear_2", F.silu(_linear(p, "add_embedding.linear_1", add)))
    return emb + aug


def _encoder_down_mid(p: Params, cfg: UNetConfig, h: Tensor, emb: Tensor, ctx: Tensor, lora=None):
    g, eps, hd = cfg.norm_groups, cfg.norm_eps, cfg.head_dim
    skips = [h]
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet_block(p, f"down_blocks.{i}.resnets.{j}", h, emb, g, eps, lora)
            if len(cfg.down_depths[i]):
                h = transformer_2d(p, f"down_blocks.{i}.attentions.{j}", h, ctx, cfg.down_depths[i][j], g, hd, lora)
            skips.append(h)
        if i < n - 1:
            h = _conv(p, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1, lora=lora)
            skips.append(h)
    h = resnet_block(p, "mid_block.resnets.0", h, emb, g, eps, lora)
    if cfg.mid_depth is not None:
        h = transformer_2d(p, "mid_block.attentions.0", h, ctx, cfg.mid_depth, g, hd, lora)
        h = resnet_block(p, "mid_block.resnets.1", h, emb, g, eps, lora)
    return h, skips


def unet_forward(p: Params, cfg: UNetConfig, x: Tensor, t: Tensor, ctx: Tensor, text_embeds: Tensor, time_ids: Tensor,
                 down_res: Optional[List[Tensor]] = None, mid_res: Optional[Tensor] = None,
                 lora: Optional[_LoRA] = None) -> Tensor:
    """UNet2DConditionModel.forward."""
    dtype = x.dtype
    emb = time_and_add_embedding(p, cfg, t, text_embeds, time_ids, dtype)
    h = _conv(p, "conv_in", x)
    h, skips = _encoder_down_mid(p, cfg, h, emb, ctx, lora)
    if down_res is not None:
        skips = [s + r for s, r in zip(skips, down_res)]
    if mid_res is not None:
        h = h + mid_res
    g, eps, hd = cfg.norm_groups, cfg.norm_eps, cfg.head_dim
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(p, f"up_blocks.{i}.resnets.{j}", h, emb, g, eps, lora)
            if len(cfg.up_depths[i]):
                h = transformer_2d(p, f"up_blocks.{i}.attentions.{j}", h, ctx, cfg.up_depths[i][j], g, hd, lora)
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(p, f"up_blocks.{i}.upsamplers.0.conv", h, lora=lora)
    h = F.silu(_gn(p, "conv_norm_out", h, g, eps))
    return _conv(p, "conv_out", h)


def controlnet_cond_embedding(p: Params, cfg: ControlNetConfig, cond: Tensor) -> Tensor:
    h = F.silu(_conv(p, "controlnet_cond_embedding.conv_in", cond))
    for i in range(len(cfg.cond_channels) - 1):
        h = F.silu(_conv(p, f"controlnet_cond_embedding.blocks.{2 * i}", h))
        h = F.silu(_conv(p, f"controlnet_cond_embedding.blocks.{2 * i + 1}", h, stride=2))
    return _conv(p, "controlnet_cond_embedding.conv_out", h)


def controlnet_forward(p: Params, cfg: ControlNetConfig, x: Tensor, t: Tensor, ctx: Tensor, text_embeds: Tensor,
                       time_ids: Tensor, cond: Tensor, conditioning_scale: float):
    """ControlNetModel.forward -> (9 down residuals, mid residual), scaled."""
    u = cfg.unet
    emb = time_and_add_embedding(p, u, t, text_embeds, time_ids, x.dtype)
    h = _conv(p, "conv_in", x) + controlnet_cond_embedding(p, cfg, cond)
    h, skips = _encoder_down_mid(p, u, h, emb, ctx)
    down = [_conv(p, f"controlnet_down_blocks.{i}", s, padding=0) * conditioning_scale for i, s in enumerate(skips)]
    mid = _conv(p, "controlnet_mid_block", h, padding=0) * conditioning_scale
    return down, mid


# ---- VAE ---------------------------------------------------------------------------------------


def _vae_attn(p: Params, pre: str, x: Tensor, groups: int, eps: float) -> Tensor:
    B, C, H, W = x.shape
    h = _gn(p, pre + ".group_norm", x, groups, eps).view(B, C, H * W).transpose(1, 2)
    q = _linear(p, pre + ".to_q", h)[:, None]
    k = _linear(p, pre + ".to_k", h)[:, None]
    v = _linear(p, pre + ".to_v", h)[:, None]
    o = F.scaled_dot_product_attention(q, k, v)[:, 0]
    o = _linear(p, pre + ".to_out.0", o)
    return o.transpose(1, 2).reshape(B, C, H, W) + x


def vae_encode_moments(p: Params, cfg: VAEConfig, x: Tensor) -> Tensor:
    """AutoencoderKL.encode up to quant_conv: [B,3,H,W] -> moments [B,2L,H/8,W/8]."""
    g, eps = cfg.norm_groups, cfg.norm_eps
    n = len(cfg.block_out_channels)
    h = _conv(p, "encoder.conv_in", x)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            h = resnet_block(p, f"encoder.down_blocks.{i}.resnets.{j}", h, None, g, eps)
        if i < n - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = _conv(p, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = resnet_block(p, "encoder.mid_block.resnets.0", h, None, g, eps)
    h = _vae_attn(p, "encoder.mid_block.attentions.0", h, g, eps)
    h = resnet_block(p, "encoder.mid_block.resnets.1", h, None, g, eps)
    h = F.silu(_gn(p, "encoder.conv_norm_out", h, g, eps))
    h = _conv(p, "encoder.conv_out", h)
    return _conv(p, "quant_conv", h, padding=0)


def vae_sample(moments: Tensor, xi: Tensor, scaling_factor: float) -> Tensor:
    """DiagonalGaussianDistribution.sample * scaling_factor (noise xi supplied by the caller)."""
    mean, logvar = moments.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    return (mean + std * xi.to(moments.dtype)) * scaling_factor


def vae_decode(p: Params, cfg: VAEConfig, z: Tensor) -> Tensor:
    """AutoencoderKL.decode: latents (already divided by scaling_factor) -> image [-1,1]."""
    g, eps = cfg.norm_groups, cfg.norm_eps
    n = len(cfg.block_out_channels)
    h = _conv(p, "post_quant_conv", z, padding=0)
    h = _conv(p, "decoder.conv_in", h)
    h = resnet_block(p, "decoder.mid_block.resnets.0", h, None, g, eps)
    h = _vae_attn(p, "decoder.mid_block.attentions.0", h, g, eps)
    h = resnet_block(p, "decoder.mid_block.resnets.1", h, None, g, eps)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(p, f"decoder.up_blocks.{i}.resnets.{j}", h, None, g, eps)
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(p, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(_gn(p, "decoder.conv_norm_out", h, g, eps))
    return _conv(p, "decoder.conv_out", h)


# ---- LCM scheduler -----------------------------------------------------------------------------


class LCMSchedule:
    """LCMScheduler restatement: scaled_linear betas, epsilon prediction,
    original_inference_steps 50, timestep_scaling 10, sigma_data 0.5, no clipping/thresholding."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 original_inference_steps=50, timestep_scaling=10.0):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        self.timestep_scaling = timestep_scaling
        self.sigma_data = 0.5
        self.timesteps: List[int] = []

    def set_timesteps(self, n: int):
        import numpy as np
        k = self.num_train_timesteps // self.original_inference_steps
        origin = np.asarray(list(range(1, self.original_inference_steps + 1))) * k - 1
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int64)
        self.timesteps = [int(v) for v in origin[idx]]
        return self.timesteps

    def img2img_timesteps(self, n: int, strength: float):
        """Pipeline get_timesteps: returns (executed timesteps, begin_index)."""
        self.set_timesteps(n)
        init = min(int(n * strength), n)
        t_start = max(n - init, 0)
        return self.timesteps[t_start:], t_start

    def add_noise_coeffs(self, t: int):
        a = float(self.alphas_cumprod[t])
        return a ** 0.5, (1.0 - a) ** 0.5

    def step_coeffs(self, step_index: int):
        """Scalars for step(eps, t, x) at global index ``step_index`` of ``self.timesteps``."""
        t = self.timesteps[step_index]
        last = step_index == len(self.timesteps) - 1
        prev_t = t if last else self.timesteps[step_index + 1]
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else 1.0
        s = t * self.timestep_scaling
        c_skip = self.sigma_data ** 2 / (s ** 2 + self.sigma_data ** 2)
        c_out = s / (s ** 2 + self.sigma_data ** 2) ** 0.5
        return dict(t=t, last=last, sqrt_a=a_t ** 0.5, sqrt_1ma=(1 - a_t) ** 0.5, c_skip=c_skip, c_out=c_out,
                    sqrt_a_prev=a_prev ** 0.5, sqrt_1ma_prev=(1 - a_prev) ** 0.5)

    def add_noise(self, z0: Tensor, noise: Tensor, t: int) -> Tensor:
        sa, s1 = self.add_noise_coeffs(t)
        # diffusers casts alphas_cumprod to the sample dtype before sqrt
        a = self.alphas_cumprod[t].to(z0.dtype)
        return a.sqrt() * z0 + (1 - a).sqrt() * noise.to(z0.dtype)

    def step(self, eps: Tensor, step_index: int, x: Tensor, noise: Optional[Tensor]) -> Tensor:
        c = self.step_coeffs(step_index)
        x0 = (x - c["sqrt_1ma"] * eps) / c["sqrt_a"]
        den = c["c_out"] * x0 + c["c_skip"] * x
        if c["last"]:
            return den
        return c["sqrt_a_prev"] * den + c["sqrt_1ma_prev"] * noise.to(x.dtype)


# ---- Pipeline ----------------------------------------------------------------------------------


def preprocess_image(img_u8: Tensor, dtype) -> Tensor:
    """VaeImageProcessor.preprocess: uint8 [B,H,W,3] -> [B,3,H,W] in [-1,1] (fp32 math, then cast)."""
    x = img_u8.permute(0, 3, 1, 2).float() / 255.0
    return (2.0 * x - 1.0).to(dtype)


def preprocess_control(edges_u8: Tensor, dtype) -> Tensor:
    """control_image_processor.preprocess (do_normalize False): uint8 [B,H,W,3] -> [B,3,H,W] in {0,1}."""
    return (edges_u8.permute(0, 3, 1, 2).float() / 255.0).to(dtype)


def postprocess_image(x: Tensor) -> Tensor:
    """VaeImageProcessor.postprocess(output_type='pil') up to the uint8 array: [B,3,H,W] -> uint8 [B,H,W,3]."""
    y = (x / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float()
    return (y * 255).round().to(torch.uint8)


@dataclass
class EditModels:
    unet_cfg: UNetConfig
    unet: Params
    cn_cfg: ControlNetConfig
    cn: Params
    vae_cfg: VAEConfig
    vae: Params
    lora: Optional[Params] = None
    lora_scale: float = 1.0


def edit_pipeline(m: EditModels, image_u8: Tensor, edges_u8: Tensor, prompt_embeds: Tensor, pooled: Tensor,
                  noises: Sequence[Tensor], strength=0.5, num_inference_steps=4, guidance_scale=1.5,
                  controlnet_conditioning_scale=0.5, dtype=torch.float32, return_all=False, on_stage=None):
    """StableDiffusionXLControlNetImg2ImgPipeline.__call__ restated.

    image_u8/edges_u8: uint8 [B,H,W,3]; prompt_embeds [2,77,D] (row 0 negative, row 1 positive), pooled [2,P]
    (shared by all images of the batch); noises: [xi, n, z1, ...] each [B,4,H/8,W/8] (RNG order of the
    reference generator: posterior sample, init noise, then one per non-final executed step).
    on_stage(name): optional callback at the start of each stage (bench.py's CPU stage split).
    """
    mark = on_stage or (lambda name: None)
    mark("vae_encode")
    dev = image_u8.device
    B, H, W, _ = image_u8.shape
    do_cfg = guidance_scale > 1
    sched = LCMSchedule()
    timesteps, begin = sched.img2img_timesteps(num_inference_steps, strength)
    x_img = preprocess_image(image_u8, dtype)
    cond = preprocess_control(edges_u8, dtype)
    moments = vae_encode_moments(m.vae, m.vae_cfg, x_img)
    z0 = vae_sample(moments, noises[0].to(dev), m.vae_cfg.scaling_factor)
    out = dict(moments=moments, z0=z0)
    if not timesteps:
        x = z0
    else:
        x = sched.add_noise(z0, noises[1].to(dev), timesteps[0])
    time_ids = torch.tensor([H, W, 0, 0, H, W], dtype=torch.float32, device=dev)
    lora = _LoRA(m.lora, m.lora_scale) if m.lora is not None else None
    nrow = 2 if do_cfg else 1
    pe = prompt_embeds if do_cfg else prompt_embeds[1:]
    pl = pooled if do_cfg else pooled[1:]
    # batch layout as diffusers: cat([neg]*B, [pos]*B)
    ctx = pe.to(dtype).repeat_interleave(B, dim=0)
    te = pl.to(dtype).repeat_interleave(B, dim=0)
    tids = time_ids[None].expand(nrow * B, -1)
    if do_cfg:
        cond = torch.cat([cond] * 2)
    zi = 2
    eps_list = []
    for k, t in enumerate(timesteps):
        x2 = torch.cat([x] * nrow)
        tt = torch.tensor([t], device=dev)
        mark("controlnet_step")
        down, mid = controlnet_forward(m.cn, m.cn_cfg, x2, tt, ctx, te, tids, cond, controlnet_conditioning_scale)
        mark("unet_step")
        eps = unet_forward(m.unet, m.unet_cfg, x2, tt, ctx, te, tids, down, mid, lora)
        if do_cfg:
            e_u, e_c = eps.chunk(2)
            eps = e_u + guidance_scale * (e_c - e_u)
        eps_list.append(eps)
        step_index = begin + k
        last = step_index == len(sched.timesteps) - 1
        z = None
        if not last:
            z = noises[zi].to(dev)
            zi += 1
        x = sched.step(eps, step_index, x, z)
    out["latents"] = x
    out["eps"] = eps_list
    mark("vae_decode")
    img = vae_decode(m.vae, m.vae_cfg, x / m.vae_cfg.scaling_factor)
    out["decoded"] = img
    out["image_u8"] = postprocess_image(img)
    mark("end")
    return out if return_all else out["image_u8"]


def ssim(a: Tensor, b: Tensor, data_range: float = 1.0) -> float:
    """SSIM as the reference defines it (torchmetrics StructuralSimilarityIndexMeasure(data_range=1.0),
    reference src/metrics.py:174-176): 11x11 Gaussian sigma 1.5, k1=0.01, k2=0.03, mean over the map.
    a, b: [B,3,H,W] in [0,1]."""
    a = a.float()
    b = b.float()
    k = torch.arange(11, dtype=torch.float32, device=a.device) - 5
    g = torch.exp(-(k ** 2) / (2 * 1.5 ** 2))
    g = (g / g.sum())
    w = (g[:, None] * g[None, :])[None, None].expand(a.shape[1], 1, 11, 11)
    C = a.shape[1]
    pad = 5
    ap = F.pad(a, (pad,) * 4, mode="reflect")
    bp = F.pad(b, (pad,) * 4, mode="reflect")
    mu_a = F.conv2d(ap, w, groups=C)
    mu_b = F.conv2d(bp, w, groups=C)
    s_aa = F.conv2d(ap * ap, w, groups=C) - mu_a ** 2
    s_bb = F.conv2d(bp * bp, w, groups=C) - mu_b ** 2
    s_ab = F.conv2d(ap * bp, w, groups=C) - mu_a * mu_b
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    m = ((2 * mu_a * mu_b + c1) * (2 * s_ab + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (s_aa + s_bb + c2))
    return float(m.mean())

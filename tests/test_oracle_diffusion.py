"""CPU: what can be pinned offline for the diffusion oracle — parameter totals of the named architectures, LoRA
fuse == unfused, fp32 tiny pipeline determinism, SSIM definition sanity, weight-packing host logic."""
import numpy as np
import torch

from oracle import diffusion_oracle as O
from fast_image_editing_with_generative_models_b200 import configs as C
from fast_image_editing_with_generative_models_b200 import synthetic as S
from fast_image_editing_with_generative_models_b200.weights import fuse_lora, pack_conv3x3


def test_parameter_totals_match_published_figures():
    with S.shapes_only():
        assert round(S.count_params(S.make_unet_params(C.sdxl_unet_config())) / 1e6, 1) == 2567.5
        assert round(S.count_params(S.make_unet_params(C.ssd1b_unet_config())) / 1e6, 1) == 1331.3
        assert round(S.count_params(S.make_controlnet_params(C.controlnet_config(True))) / 1e6, 1) == 1251.0
        assert round(S.count_params(S.make_vae_params(C.VAEConfig())) / 1e6, 1) == 83.7


def test_lora_fuse_equals_unfused_reference_path():
    torch.manual_seed(0)
    ucfg = C.tiny_unet_config()
    up = S.make_unet_params(ucfg)
    lp = S.make_lora_params(up, rank=8)
    fused = dict(up)
    for k in list(lp):
        if k.endswith(".lora_A.weight"):
            base = k[: -len(".lora_A.weight")]
            fused[base + ".weight"] = fuse_lora(up[base + ".weight"], lp[k], lp[base + ".lora_B.weight"], 1.0)
    x = torch.randn(2, 4, 16, 16); ctx = torch.randn(2, 77, ucfg.cross_attention_dim); te = torch.randn(2, 64)
    tids = torch.tensor([[128.0, 128, 0, 0, 128, 128]] * 2)
    t = torch.tensor([499])
    a = O.unet_forward(up, ucfg, x, t, ctx, te, tids, lora=O._LoRA(lp, 1.0))
    b = O.unet_forward(fused, ucfg, x, t, ctx, te, tids)
    assert float((a - b).abs().max()) < 1e-4
    assert float((a - O.unet_forward(up, ucfg, x, t, ctx, te, tids)).abs().max()) > 1e-4   # the LoRA path is non-trivial


def test_tiny_pipeline_runs_and_is_deterministic():
    ucfg, ccfg, vcfg = C.tiny_unet_config(), C.tiny_controlnet_config(), C.tiny_vae_config()
    m = O.EditModels(ucfg, S.make_unet_params(ucfg), ccfg, S.make_controlnet_params(ccfg), vcfg, S.make_vae_params(vcfg))
    img = torch.from_numpy(S.synthetic_image(0, 64, 64)[None])
    from oracle.canny_oracle import preprocess_image
    edges = torch.from_numpy(preprocess_image(img[0].numpy())[None])
    pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, 64)
    nz = S.synthetic_noises(0, 1, 8, 8)
    a = O.edit_pipeline(m, img, edges, pe.float(), pl.float(), nz, return_all=True)
    b = O.edit_pipeline(m, img, edges, pe.float(), pl.float(), nz, return_all=True)
    assert torch.equal(a["image_u8"], b["image_u8"]) and a["image_u8"].shape == (1, 64, 64, 3)
    assert len(a["eps"]) == 2                       # strength 0.5 of 4 steps executes 2 evaluations
    c = O.edit_pipeline(m, img, edges, pe.float(), pl.float(), nz, strength=0.8, return_all=True)
    assert len(c["eps"]) == 3


def test_ssim_definition():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 3, 64, 64, generator=g)
    assert abs(O.ssim(a, a) - 1.0) < 1e-6
    assert O.ssim(a, torch.rand(1, 3, 64, 64, generator=g)) < 0.2


def test_pack_conv3x3_layout():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv3x3(w, pad_cout_to=4, pad_cin_to=8)
    assert p.shape == (4, 72) and p.dtype == torch.float16
    assert float(p[1].reshape(3, 3, 8)[2, 1, 2]) == float(w[1, 2, 2, 1]) and float(p[2:].abs().max()) == 0 and float(p[:, 3:8].abs().max()) == 0


def test_state_dict_key_counts_and_exact_totals_match_public_checkpoints():
    """Public facts about the named Hugging Face checkpoints (recalled, not fetched — there is no network here): the diffusers state
    dict of stabilityai/stable-diffusion-xl-base-1.0 `unet` has 1680 tensors / 2,567,463,684 parameters; the SDXL AutoencoderKL
    (both madebyollin/sdxl-vae-fp16-fix and stabilityai/sdxl-vae) has 248 tensors / 83,653,863 parameters; latent-consistency/
    lcm-lora-sdxl adapts 788 modules at rank 64 (197 M parameters; 788 x {down, up, alpha} = 2364 tensors in the kohya file).
    A restatement that walked a different topology (a missing attention, a wrong shortcut, the wrong LoRA target set) would
    miss these exact integers."""
    with S.shapes_only():
        unet = S.make_unet_params(C.sdxl_unet_config())
        assert (len(unet), S.count_params(unet)) == (1680, 2_567_463_684)
        vae = S.make_vae_params(C.VAEConfig())
        assert (len(vae), S.count_params(vae)) == (248, 83_653_863)
        lora = S.make_lora_params(unet, rank=64)
        assert len(lora) == 2 * 788 and round(S.count_params(lora) / 1e6) == 197
        # every key carries the diffusers naming scheme the real checkpoints use (spot checks across block types)
        for k in ("down_blocks.1.attentions.0.transformer_blocks.1.attn2.to_k.weight", "mid_block.attentions.0.transformer_blocks.9.ff.net.0.proj.bias",
                  "up_blocks.0.resnets.2.conv_shortcut.weight", "up_blocks.1.upsamplers.0.conv.weight", "add_embedding.linear_1.weight",
                  "time_embedding.linear_2.bias", "conv_norm_out.weight", "down_blocks.0.downsamplers.0.conv.bias"):
            assert k in unet, k
        assert tuple(unet["add_embedding.linear_1.weight"].shape) == (1280, 2816) and tuple(unet["up_blocks.0.resnets.0.conv1.weight"].shape) == (1280, 2560, 3, 3)
        for k in ("encoder.mid_block.attentions.0.to_q.weight", "decoder.up_blocks.3.resnets.2.conv2.weight", "quant_conv.weight", "post_quant_conv.bias",
                  "decoder.up_blocks.2.resnets.0.conv_shortcut.weight"):
            assert k in vae, k
        ssd = S.make_unet_params(C.ssd1b_unet_config())
        assert "mid_block.attentions.0.proj_in.weight" not in ssd and "mid_block.resnets.1.conv1.weight" not in ssd      # SSD-1B: attention-free single-resnet mid block
        assert "mid_block.resnets.0.conv1.weight" in ssd


def test_blocks_match_an_independent_torch_nn_restatement():
    """Second, independent restatement of the two block types that carry the conventions most easily got wrong — head split and
    q/k/v packing (nn.MultiheadAttention owns its own reshape), GEGLU operand order, where the time embedding enters a resnet —
    built from torch.nn MODULES and loaded with the oracle's diffusers-named parameters; the functional oracle must agree."""
    import torch.nn as nn
    torch.manual_seed(1)
    ucfg = C.tiny_unet_config()
    p = S.make_unet_params(ucfg)
    pre = next(k for k in p if k.endswith("transformer_blocks.0.attn1.to_q.weight"))[: -len(".attn1.to_q.weight")]
    c = p[pre + ".attn1.to_q.weight"].shape[0]
    dctx = p[pre + ".attn2.to_k.weight"].shape[1]
    x, ctx = torch.randn(2, 48, c), torch.randn(2, 77, dctx)

    class Block(nn.Module):
        def __init__(s):
            super().__init__()
            s.n1, s.n2, s.n3 = nn.LayerNorm(c), nn.LayerNorm(c), nn.LayerNorm(c)
            s.a1 = nn.MultiheadAttention(c, c // 64, bias=False, batch_first=True)
            s.a2 = nn.MultiheadAttention(c, c // 64, bias=False, batch_first=True, kdim=dctx, vdim=dctx)
            s.o1, s.o2 = nn.Linear(c, c), nn.Linear(c, c)
            s.ff_in, s.ff_out = nn.Linear(c, 8 * c), nn.Linear(4 * c, c)

        def forward(s, x, ctx):
            h = s.n1(x)
            x = x + s.o1(s.a1(h, h, h, need_weights=False)[0])
            x = x + s.o2(s.a2(s.n2(x), ctx, ctx, need_weights=False)[0])
            val, gate = s.ff_in(s.n3(x)).split(4 * c, dim=-1)
            return x + s.ff_out(val * nn.functional.gelu(gate))

    b = Block()
    ident = torch.eye(c)
    with torch.no_grad():
        for i, n in enumerate(("n1", "n2", "n3"), 1):
            getattr(b, n).weight.copy_(p[f"{pre}.norm{i}.weight"]); getattr(b, n).bias.copy_(p[f"{pre}.norm{i}.bias"])
        b.a1.in_proj_weight.copy_(torch.cat([p[f"{pre}.attn1.to_{t}.weight"] for t in "qkv"]))
        b.a1.out_proj.weight.copy_(ident)                                      # diffusers' to_out.0 is applied separately (it has a bias)
        if b.a2.in_proj_weight is not None:                                    # kdim == embed_dim: torch packs q, k, v
            b.a2.in_proj_weight.copy_(torch.cat([p[f"{pre}.attn2.to_{t}.weight"] for t in "qkv"]))
        else:
            b.a2.q_proj_weight.copy_(p[f"{pre}.attn2.to_q.weight"]); b.a2.k_proj_weight.copy_(p[f"{pre}.attn2.to_k.weight"]); b.a2.v_proj_weight.copy_(p[f"{pre}.attn2.to_v.weight"])
        b.a2.out_proj.weight.copy_(ident)
        for mod, name in ((b.o1, "attn1.to_out.0"), (b.o2, "attn2.to_out.0"), (b.ff_in, "ff.net.0.proj"), (b.ff_out, "ff.net.2")):
            mod.weight.copy_(p[f"{pre}.{name}.weight"]); mod.bias.copy_(p[f"{pre}.{name}.bias"])
        want = b(x, ctx)
        got = O.transformer_block(p, pre, x, ctx, 64)
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max())

    rp = next(k for k in p if k.endswith("resnets.0.conv_shortcut.weight"))[: -len(".conv_shortcut.weight")]
    cout, cin = p[rp + ".conv1.weight"].shape[:2]
    tdim = p[rp + ".time_emb_proj.weight"].shape[1]

    class Res(nn.Module):
        def __init__(s):
            super().__init__()
            s.norm1, s.conv1 = nn.GroupNorm(ucfg.norm_groups, cin, eps=ucfg.norm_eps), nn.Conv2d(cin, cout, 3, padding=1)
            s.time_emb_proj = nn.Linear(tdim, cout)
            s.norm2, s.conv2 = nn.GroupNorm(ucfg.norm_groups, cout, eps=ucfg.norm_eps), nn.Conv2d(cout, cout, 3, padding=1)
            s.conv_shortcut = nn.Conv2d(cin, cout, 1)
            s.act = nn.SiLU()

        def forward(s, x, temb):
            h = s.conv1(s.act(s.norm1(x))) + s.time_emb_proj(s.act(temb))[:, :, None, None]
            return s.conv_shortcut(x) + s.conv2(s.act(s.norm2(h)))

    r = Res()
    r.load_state_dict({k[len(rp) + 1:]: v for k, v in p.items() if k.startswith(rp + ".")})       # the oracle's names ARE the module names
    xi, temb = torch.randn(2, cin, 8, 8), torch.randn(2, tdim)
    with torch.no_grad():
        want = r(xi, temb)
        got = O.resnet_block(p, rp, xi, temb, ucfg.norm_groups, ucfg.norm_eps)
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max())


def test_vae_attention_matches_an_independent_torch_nn_restatement():
    """AutoencoderKL mid-block attention: ONE head over all C channels (scale 1 / sqrt(C)), GroupNorm before, residual after — against
    nn.GroupNorm + nn.MultiheadAttention(num_heads=1), which own their reshapes and their scale."""
    import torch.nn as nn
    torch.manual_seed(2)
    vcfg = C.tiny_vae_config()
    p = S.make_vae_params(vcfg)
    pre = "encoder.mid_block.attentions.0"
    c = p[pre + ".to_q.weight"].shape[0]
    gn = nn.GroupNorm(vcfg.norm_groups, c, eps=vcfg.norm_eps)
    mha = nn.MultiheadAttention(c, 1, bias=True, batch_first=True)
    with torch.no_grad():
        gn.weight.copy_(p[pre + ".group_norm.weight"]); gn.bias.copy_(p[pre + ".group_norm.bias"])
        mha.in_proj_weight.copy_(torch.cat([p[f"{pre}.to_{t}.weight"] for t in "qkv"]))
        mha.in_proj_bias.copy_(torch.cat([p[f"{pre}.to_{t}.bias"] for t in "qkv"]))
        mha.out_proj.weight.copy_(p[pre + ".to_out.0.weight"]); mha.out_proj.bias.copy_(p[pre + ".to_out.0.bias"])
        x = torch.randn(2, c, 6, 5)
        h = gn(x).flatten(2).transpose(1, 2)
        want = mha(h, h, h, need_weights=False)[0].transpose(1, 2).reshape(x.shape) + x
        got = O._vae_attn(p, pre, x, vcfg.norm_groups, vcfg.norm_eps)
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max())

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

"""GPU: full-size (1024x1024) parity of the complete edit path against the fp32 oracle with identical seeded
random-init weights of the named architectures — the BASELINE.json acceptance criteria:
Canny bit-exact, final latents max-abs <= 2e-2, decoded SSIM >= 0.99."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import diffusion_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,batch", [("ssd-1b", 1), ("sdxl", 2)])
def test_full_edit_parity(cuda_dev, model, batch):
    from fast_image_editing_with_generative_models_b200 import model_zoo
    from fast_image_editing_with_generative_models_b200 import synthetic as S
    state = model_zoo.synthetic_state(model)
    eng = model_zoo.build_engine(state, cuda_dev)
    ucfg = state["unet_cfg"]
    H = 1024
    imgs = np.stack([S.synthetic_image(s, H, H) for s in range(batch)])
    pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim)
    noises = S.synthetic_noises(0, batch, H // 8, H // 8)
    d_img = torch.from_numpy(imgs).to(cuda_dev)
    out = eng.edit_batch(d_img, pe, pl, noises, strength=0.5, return_extras=True)
    edges_ref = c_oracle.canny_u8(imgs, 100, 200, replicate3=True)
    assert np.array_equal(out.edges.cpu().numpy(), edges_ref), "Canny control image not bit-exact"
    del eng
    torch.cuda.empty_cache()
    f32 = lambda p: None if p is None else O.to_dtype(p, torch.float32, cuda_dev)
    m = O.EditModels(ucfg, f32(state["unet"]), state["cn_cfg"], f32(state["cn"]), state["vae_cfg"], f32(state["vae"]), f32(state["lora"]), 1.0)
    with torch.no_grad():
        ref = O.edit_pipeline(m, d_img, torch.from_numpy(edges_ref).to(cuda_dev), pe.float().to(cuda_dev), pl.float().to(cuda_dev), noises,
                              strength=0.5, dtype=torch.float32, return_all=True)
    nchw = lambda x: x.permute(0, 3, 1, 2).float()
    mom_err = float((nchw(out.extras["moments"]) - ref["moments"]).abs().max())
    eps_err = [float((nchw(a)[: batch] - 0).abs().max()) for a in out.extras["eps"]]
    lat_err = float((nchw(out.latents) - ref["latents"]).abs().max())
    lat_ref_max = float(ref["latents"].abs().max())
    a = out.images.permute(0, 3, 1, 2).float() / 255.0
    b = ref["image_u8"].permute(0, 3, 1, 2).float() / 255.0
    s = O.ssim(a, b)
    px = float((out.images.float() - ref["image_u8"].float()).abs().mean())
    print(f"\n[{model} b{batch}] moments max-abs {mom_err:.4g}; latents max-abs {lat_err:.4g} (ref absmax {lat_ref_max:.3g}, rel {lat_err / lat_ref_max:.3g}); "
          f"eps std {[float(e.std()) for e in ref['eps']]}; decoded std {float(ref['decoded'].std()):.3f} mean {float(ref['decoded'].mean()):.3f}; "
          f"SSIM {s:.5f}; mean |dpx| {px:.3f}; engine eps absmax {eps_err}")
    # noise floor: the oracle's own fp16-vs-fp32 gap (torch fp16 ops = what the reference would run on a GPU)
    m16 = O.EditModels(ucfg, O.to_dtype(state["unet"], torch.float16, cuda_dev), state["cn_cfg"], O.to_dtype(state["cn"], torch.float16, cuda_dev),
                       state["vae_cfg"], O.to_dtype(state["vae"], torch.float16, cuda_dev),
                       None if state["lora"] is None else O.to_dtype(state["lora"], torch.float16, cuda_dev), 1.0)
    with torch.no_grad():
        ref16 = O.edit_pipeline(m16, d_img, torch.from_numpy(edges_ref).to(cuda_dev), pe.to(cuda_dev), pl.to(cuda_dev), noises,
                                strength=0.5, dtype=torch.float16, return_all=True)
    floor = float((ref16["latents"].float() - ref["latents"]).abs().max())
    print(f"[{model} b{batch}] torch-fp16 oracle vs fp32 oracle: latents max-abs {floor:.4g}; SSIM {O.ssim(ref16['image_u8'].permute(0, 3, 1, 2).float() / 255.0, b):.5f}")
    assert lat_err <= 2e-2, lat_err
    assert s >= 0.99, s

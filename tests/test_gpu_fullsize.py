"""GPU: full-size (1024x1024) parity of the complete edit path against the fp32 oracle with identical seeded
random-init weights of the named architectures — the BASELINE.json acceptance criteria:
Canny bit-exact, final latents max-abs <= 2e-2, decoded SSIM >= 0.99.

Cases: the BASELINE configurations (strength 0.5: SSD-1B batch 1, SDXL batch 2), the reference's actual CLI default
(strength 0.80 -> 3 executed steps ``[759, 499, 259]``, two step-noise draws; ``/root/reference/src/pipeline.py:217``,
``run_batch.py:209-219``), ``use_full_controlnet=True`` (``src/pipeline.py:82-84``), and the benchmarked shape itself
(SDXL, 8 images, CUDA-graph replay) with graph == eager bit for bit."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import diffusion_oracle as O

pytestmark = pytest.mark.gpu

_STATES = {}


def _state(model, full_cn=False):
    from fast_image_editing_with_generative_models_b200 import model_zoo
    key = (model, full_cn)
    if key not in _STATES:
        _STATES.clear()                      # one set of fp32 master weights (10+ GB for SDXL) at a time
        _STATES[key] = model_zoo.synthetic_state(model, full_cn)
    return _STATES[key]


def _inputs(state, batch, n_noise=4):
    from fast_image_editing_with_generative_models_b200 import synthetic as S
    ucfg = state["unet_cfg"]
    H = 1024
    imgs = np.stack([S.synthetic_image(s, H, H) for s in range(batch)])
    pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim)
    noises = S.synthetic_noises(0, batch, H // 8, H // 8, n_noise)
    return imgs, pe, pl, noises


def _oracle(state, dev, dtype, d_img, edges_ref, pe, pl, noises, strength):
    cast = lambda p: None if p is None else O.to_dtype(p, dtype, dev)
    m = O.EditModels(state["unet_cfg"], cast(state["unet"]), state["cn_cfg"], cast(state["cn"]), state["vae_cfg"], cast(state["vae"]), cast(state["lora"]), 1.0)
    with torch.no_grad():
        return O.edit_pipeline(m, d_img, torch.from_numpy(edges_ref).to(dev), pe.to(dev, dtype), pl.to(dev, dtype), noises, strength=strength,
                               dtype=dtype, return_all=True)


def _compare(tag, out_latents, out_images, ref):
    nchw = lambda x: x.permute(0, 3, 1, 2).float()
    lat_err = float((nchw(out_latents) - ref["latents"]).abs().max())
    lat_ref_max = float(ref["latents"].abs().max())
    a = out_images.permute(0, 3, 1, 2).float() / 255.0
    b = ref["image_u8"].permute(0, 3, 1, 2).float() / 255.0
    per_image = [O.ssim(a[i:i + 1], b[i:i + 1]) for i in range(a.shape[0])]
    px = float((out_images.float() - ref["image_u8"].float()).abs().mean())
    print(f"\n[{tag}] latents max-abs {lat_err:.4g} (ref absmax {lat_ref_max:.3g}, rel {lat_err / lat_ref_max:.3g}); eps std {[round(float(e.std()), 3) for e in ref['eps']]}; "
          f"decoded std {float(ref['decoded'].std()):.3f}; SSIM min {min(per_image):.5f}; mean |dpx| {px:.3f}")
    return lat_err, min(per_image), lat_ref_max


def _latent_tolerance(ref_absmax):
    """BASELINE.json states max-abs 2e-2 for its configuration (strength 0.5), where the reference latents reach |x| = 14.5, i.e. a
    relative accuracy of 1.4e-3.  With more executed steps the synthetic-weight latents grow (|x| = 40 at strength 0.8), so the
    criterion is applied at the same relative accuracy: 2e-2 * max(1, |x|max / 14.5)."""
    return 2e-2 * max(1.0, ref_absmax / 14.5)


@pytest.mark.parametrize("model,batch,strength,full_cn,floor", [("ssd-1b", 1, 0.5, False, True), ("ssd-1b", 1, 0.8, False, True), ("ssd-1b", 1, 0.5, True, False),
                                                               ("sdxl", 2, 0.5, False, True), ("sdxl", 1, 0.8, False, False)])
def test_full_edit_parity(cuda_dev, model, batch, strength, full_cn, floor):
    from fast_image_editing_with_generative_models_b200 import model_zoo
    state = _state(model, full_cn)
    eng = model_zoo.build_engine(state, cuda_dev)
    imgs, pe, pl, noises = _inputs(state, batch)
    d_img = torch.from_numpy(imgs).to(cuda_dev)
    out = eng.edit_batch(d_img, pe, pl, noises, strength=strength, return_extras=True)
    n_exec = int(4 * strength)
    assert len(out.extras["eps"]) == n_exec                                   # strength 0.8 runs [759, 499, 259]
    edges_ref = c_oracle.canny_u8(imgs, 100, 200, replicate3=True)
    assert np.array_equal(out.edges.cpu().numpy(), edges_ref), "Canny control image not bit-exact"
    mom = out.extras["moments"]
    latents, images = out.latents.clone(), out.images.clone()
    del eng, out
    torch.cuda.empty_cache()
    ref = _oracle(state, cuda_dev, torch.float32, d_img, edges_ref, pe, pl, noises, strength)
    mom_err = float((mom.permute(0, 3, 1, 2).float() - ref["moments"]).abs().max())
    tag = f"{model} b{batch} strength {strength}{' full-controlnet' if full_cn else ''}"
    lat_err, s, ref_max = _compare(tag, latents, images, ref)
    print(f"[{tag}] moments max-abs {mom_err:.4g}")
    if floor:
        # noise floor: the oracle's own fp16-vs-fp32 gap (torch fp16 ops = what the reference would run on a GPU)
        ref16 = _oracle(state, cuda_dev, torch.float16, d_img, edges_ref, pe, pl, noises, strength)
        fl = float((ref16["latents"].float() - ref["latents"]).abs().max())
        print(f"[{tag}] torch-fp16 oracle vs fp32 oracle: latents max-abs {fl:.4g}; "
              f"SSIM {O.ssim(ref16['image_u8'].permute(0, 3, 1, 2).float() / 255.0, ref['image_u8'].permute(0, 3, 1, 2).float() / 255.0):.5f}")
        assert lat_err <= fl, (lat_err, fl)          # closer to the fp32 oracle than torch's own fp16 execution of the same modules
    if strength == 0.5:
        assert lat_err <= 2e-2, lat_err               # the BASELINE.json criterion, verbatim, on its own configuration
    assert lat_err <= _latent_tolerance(ref_max), (lat_err, ref_max)
    assert s >= 0.99, s


def test_bench_shape_sdxl_b8_graph_parity(cuda_dev):
    """The shape bench.py times: SDXL, 8 images per GPU, strength 0.5, replayed as one CUDA graph.  Every image must meet the
    criteria, and the graph replay must equal the eager launch sequence bit for bit."""
    from fast_image_editing_with_generative_models_b200 import model_zoo
    state = _state("sdxl")
    eng = model_zoo.build_engine(state, cuda_dev)
    B = 8
    imgs, pe, pl, noises = _inputs(state, B)
    d_img = torch.from_numpy(imgs).to(cuda_dev)
    eager = eng.edit_batch(d_img, pe, pl, noises, strength=0.5, return_latents=True, use_graph=False)
    e_lat, e_img = eager.latents.clone(), eager.images.clone()
    g1 = eng.edit_batch(d_img, pe, pl, noises, strength=0.5, return_latents=True, use_graph=True)       # capture + first replay
    g1_lat = g1.latents.clone()
    g2 = eng.edit_batch(d_img, pe, pl, noises, strength=0.5, return_latents=True, use_graph=True)       # pure replay
    assert torch.equal(g1_lat, e_lat) and torch.equal(g2.latents, e_lat) and torch.equal(g2.images, e_img), "graph replay differs from eager"
    edges_ref = c_oracle.canny_u8(imgs, 100, 200, replicate3=True)
    assert np.array_equal(g2.edges.cpu().numpy(), edges_ref)
    latents, images = g2.latents.clone(), g2.images.clone()
    del eng, eager, g1, g2
    torch.cuda.empty_cache()
    # the images of a batch are independent: the oracle runs them two at a time (fp32 activations of 8 x 1024^2 are ~5 GB per tensor)
    worst_lat, worst_ssim = 0.0, 1.0
    for i in range(0, B, 2):
        ref = _oracle(state, cuda_dev, torch.float32, d_img[i:i + 2], edges_ref[i:i + 2], pe, pl, [n[i:i + 2] for n in noises], 0.5)
        le, s, _ = _compare(f"sdxl b8 graph, images {i}-{i + 1}", latents[i:i + 2], images[i:i + 2], ref)
        worst_lat, worst_ssim = max(worst_lat, le), min(worst_ssim, s)
        del ref
    assert worst_lat <= 2e-2, worst_lat
    assert worst_ssim >= 0.99, worst_ssim

"""GPU: the evaluation metrics (reference src/metrics.py) on the fie_b200 kernels against oracle/metrics_oracle.py and the
real third-party modules this image has (transformers' CLIP, torchvision's SqueezeNet, Pillow, torch's antialiased interpolate), with the
identical seeded random-init weights.  Integer / byte work is bit-exact; floating-point tolerances are written at each assert."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def _img(seed, h=512, w=512, smooth=True):
    rng = np.random.RandomState(seed)
    x = rng.rand(h, w, 3)
    if smooth:
        for _ in range(3):
            x = (x + np.roll(x, 1, 0) + np.roll(x, 1, 1) + np.roll(x, -1, 0)) / 4
        x = (x - x.min()) / (x.max() - x.min())
    return (x * 255).astype(np.uint8)


def _noisy(a, seed, amp):
    return np.clip(a.astype(int) + np.random.RandomState(seed).randint(-amp, amp + 1, a.shape), 0, 255).astype(np.uint8)


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---- kernels ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w", [(512, 512), (97, 83), (11, 11), (43, 300)])
def test_ssim_kernel(cuda_dev, h, w):
    from fast_image_editing_with_generative_models_b200 import ops
    a = np.stack([_img(1, h, w), _img(2, h, w), np.full((h, w, 3), 77, np.uint8)])
    b = np.stack([_noisy(a[0], 3, 12), _img(4, h, w), np.full((h, w, 3), 77, np.uint8)])
    got = ops.ssim_u8(_dev(a, cuda_dev), _dev(b, cuda_dev)).cpu().numpy()
    for i in range(3):
        ref64, ref32 = MO.ssim(a[i], b[i], dtype=torch.float64), MO.ssim(a[i], b[i])
        assert abs(got[i] - ref64) < 1e-9, (i, got[i], ref64)            # double moments on the reference's float32 pixels
        assert abs(got[i] - ref32) < 1e-3, (i, got[i], ref32)            # torchmetrics' own float32 arithmetic (1e-4 of noise in flat regions)
    assert abs(got[2] - 1.0) < 1e-12


def test_sqdiff_u8_is_exact(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    a = np.stack([_img(5, 301, 200, False), _img(6, 301, 200, False), _img(7, 301, 200, False)])
    b = np.stack([_img(8, 301, 200, False), a[1], 255 - a[2]])
    got = ops.sqdiff_u8(_dev(a, cuda_dev), _dev(b, cuda_dev)).cpu().numpy()
    ref = ((a.astype(np.int64) - b.astype(np.int64)) ** 2).reshape(3, -1).sum(1)
    assert np.array_equal(got, ref) and got[1] == 0
    # plane size a multiple of 16 bytes: the 128-bit / dp4a path, incl. the largest possible differences
    a = np.stack([_img(9, 64, 80, False), np.zeros((64, 80, 3), np.uint8), _img(10, 64, 80, False)])
    b = np.stack([_img(11, 64, 80, False), np.full((64, 80, 3), 255, np.uint8), a[2]])
    got = ops.sqdiff_u8(_dev(a, cuda_dev), _dev(b, cuda_dev)).cpu().numpy()
    ref = ((a.astype(np.int64) - b.astype(np.int64)) ** 2).reshape(3, -1).sum(1)
    assert np.array_equal(got, ref) and got[1] == 255 * 255 * 64 * 80 * 3 and got[2] == 0


@pytest.mark.parametrize("h,w,oh,ow", [(512, 512, 224, 224), (1024, 1024, 224, 224), (300, 200, 336, 224), (100, 120, 224, 268)])
def test_bicubic_resize_bit_identical_to_pillow(cuda_dev, h, w, oh, ow):
    from PIL import Image
    from fast_image_editing_with_generative_models_b200 import ops
    imgs = np.stack([_img(h + i, h, w, smooth=bool(i)) for i in range(2)])
    got = ops.resize_pillow(_dev(imgs, cuda_dev), oh, ow, "bicubic").cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], np.asarray(Image.fromarray(imgs[i]).resize((ow, oh), Image.BICUBIC)))


@pytest.mark.parametrize("h,w,oh,ow", [(512, 512, 224, 224), (1024, 1024, 224, 224), (300, 400, 224, 298), (224, 224, 224, 224), (224, 300, 224, 224)])
def test_antialiased_resize_normalize_matches_torch(cuda_dev, h, w, oh, ow):
    from fast_image_editing_with_generative_models_b200 import ops
    imgs = np.stack([_img(h + w + i, h, w, smooth=bool(i)) for i in range(2)])
    got = ops.resize_aa_normalize(_dev(imgs, cuda_dev), oh, ow, MO.IMAGENET_MEAN, MO.IMAGENET_STD).cpu()
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0
    if (oh, ow) != (h, w):
        x = F.interpolate(x, size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)
    ref = ((x - torch.tensor(MO.IMAGENET_MEAN).view(1, 3, 1, 1)) / torch.tensor(MO.IMAGENET_STD).view(1, 3, 1, 1)).permute(0, 2, 3, 1)
    assert float((got - ref).abs().max()) < 5e-6                      # fp32 both; summation order differs
    # fp32 input form
    xf = torch.rand(1, h, w, 3, generator=torch.Generator().manual_seed(0))
    got = ops.resize_aa_normalize(xf.to(cuda_dev), oh, ow).cpu()
    ref = xf.permute(0, 3, 1, 2)
    if (oh, ow) != (h, w):
        ref = F.interpolate(ref, size=(oh, ow), mode="bilinear", antialias=True, align_corners=False)
    assert float((got - ref.permute(0, 2, 3, 1)).abs().max()) < 2e-6


def test_patchify_assemble_l2norm_cosine_sqdiff(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    g = torch.Generator().manual_seed(1)
    img = _img(9, 64, 48, False)[None].repeat(2, 0)
    img[1] = _img(10, 64, 48, False)
    rows = ops.patchify(_dev(img, cuda_dev), 16, MO.CLIP_MEAN, MO.CLIP_STD).float().cpu()
    x = (torch.from_numpy(img).float() / 255.0 - torch.tensor(MO.CLIP_MEAN)) / torch.tensor(MO.CLIP_STD)           # [2,64,48,3]
    ref = x.view(2, 4, 16, 3, 16, 3).permute(0, 1, 3, 2, 4, 5).reshape(2 * 12, 16 * 16 * 3)
    assert float((rows - ref).abs().max()) < 2e-3                     # fp16 rounding of values up to 2.7
    xf = torch.rand(1, 32, 32, 3, generator=g)
    rows = ops.patchify(xf.to(cuda_dev), 8).float().cpu()
    assert float((rows - xf.view(1, 4, 8, 4, 8, 3).permute(0, 1, 3, 2, 4, 5).reshape(16, 192)).abs().max()) < 5e-4
    # tokens
    patches, cls, pos = torch.randn(2 * 12, 128, generator=g).half(), torch.randn(128, generator=g).half(), torch.randn(13, 128, generator=g).half()
    tok = ops.vit_assemble(patches.to(cuda_dev), cls.to(cuda_dev), pos.to(cuda_dev), 2).float().cpu().view(2, 13, 128)
    ref = torch.cat([cls.float().expand(2, 1, 128), patches.float().view(2, 12, 128)], 1) + pos.float()
    assert float((tok - ref).abs().max()) < 4e-3
    # rows / max(|row|, eps) on a strided view
    big = torch.randn(37, 3 * 96, generator=g).half().to(cuda_dev)
    kn = ops.l2norm_rows(big[:, 96:192]).float().cpu()
    ref = F.normalize(big[:, 96:192].float().cpu(), dim=1)
    assert float((kn - ref).abs().max()) < 6e-4
    # cosine of row pairs
    a, b = torch.randn(5, 512, generator=g).half(), torch.randn(5, 512, generator=g).half()
    cs = ops.cosine_rows(a.to(cuda_dev), b.to(cuda_dev)).cpu()
    assert float((cs - F.cosine_similarity(a.float(), b.float(), dim=1)).abs().max()) < 1e-6
    # mean squared difference of fp32 windows with row strides
    m0, m1 = torch.randn(50, 64, generator=g), torch.randn(50, 96, generator=g)
    s = float(ops.sqdiff_f32(m0.to(cuda_dev), m1.to(cuda_dev), 45, 50).item())
    assert abs(s - float(((m0[:45, :50].double() - m1[:45, :50].double()) ** 2).sum())) < 1e-9 * s


def test_lpips_helper_kernels(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 31, 29, 64, generator=g).half()
    xd = x.to(cuda_dev)
    for stride, pad in ((1, 1), (2, 0)):
        cols = ops.im2col3x3(xd, stride, pad).float().cpu()
        ref = F.unfold(x.float().permute(0, 3, 1, 2), 3, padding=pad, stride=stride)                       # [2, C*9, L], rows (c, ky, kx)
        ref = ref.view(2, 64, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * 64)                               # -> [(n, oy, ox), (ky, kx, c)]
        assert torch.equal(cols, ref)
    # channel-strided view (the 16 squeeze channels of a 64-wide buffer) and K padding
    cols = ops.im2col3x3(xd[..., :16], 1, 1, kpad=192).float().cpu()
    ref = F.unfold(x[..., :16].float().permute(0, 3, 1, 2), 3, padding=1).view(2, 16, 9, -1).permute(0, 3, 2, 1).reshape(-1, 144)
    assert torch.equal(cols[:, :144], ref) and float(cols[:, 144:].abs().max()) == 0.0
    # the uint8 network input with the LPIPS scaling layer
    img = _img(11, 33, 35, False)[None]
    cols = ops.im2col3x3(_dev(img, cuda_dev), 2, 0, MO.LPIPS_SHIFT, MO.LPIPS_SCALE, kpad=64).float().cpu()
    xin = ((torch.from_numpy(img).float() / 255.0 * 2 - 1) - torch.tensor(MO.LPIPS_SHIFT)) / torch.tensor(MO.LPIPS_SCALE)
    ref = F.unfold(xin.permute(0, 3, 1, 2), 3, stride=2).view(1, 3, 9, -1).permute(0, 3, 2, 1).reshape(-1, 27)
    assert float((cols[:, :27] - ref).abs().max()) < 2e-3 and float(cols[:, 27:].abs().max()) == 0.0
    # ceil-mode pooling on odd sizes
    for h, w in ((255, 255), (127, 63), (31, 32), (4, 3)):
        t = torch.randn(2, h, w, 24, generator=g).half()
        got = ops.maxpool3s2_ceil(t.to(cuda_dev)).cpu()
        ref = F.max_pool2d(t.float().permute(0, 3, 1, 2), 3, 2, ceil_mode=True).permute(0, 2, 3, 1).half()
        assert got.shape == ref.shape and torch.equal(got, ref), (h, w)
    # one LPIPS layer
    f0, f1 = torch.randn(2, 9, 7, 384, generator=g).relu().half(), torch.randn(2, 9, 7, 384, generator=g).relu().half()
    f1[1] = f0[1]
    lin = torch.rand(384, generator=g)
    got = ops.lpips_layer(f0.to(cuda_dev), f1.to(cuda_dev), lin.to(cuda_dev)).cpu()
    n0 = f0.double() / (f0.double().pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    n1 = f1.double() / (f1.double().pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    ref = (((n0 - n1) ** 2) * lin.double()).sum(-1).mean((1, 2))
    assert float((got - ref).abs().max()) < 1e-5 * float(ref.max()) + 1e-12 and float(got[1]) == 0.0      # fp32 unit vectors, double sums


def test_gemm_relu_epilogue_into_concat_buffer(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    g = torch.Generator().manual_seed(3)
    m = 255 * 3 + 5
    a, w1, w3 = torch.randn(m, 64, generator=g).half(), (torch.randn(64, 64, generator=g) / 8).half(), (torch.randn(192, 576, generator=g) / 24).half()
    b1, b3 = torch.randn(64, generator=g), torch.randn(192, generator=g)
    a3 = torch.randn(m, 576, generator=g).half()
    out = torch.full((m, 256), 7.0, dtype=torch.float16, device=cuda_dev)
    ops.gemm(a.to(cuda_dev), w1.to(cuda_dev), col_bias=b1.to(cuda_dev), act=ops.ACT_RELU, out=out[:, :64])
    ops.gemm(a3.to(cuda_dev), w3.to(cuda_dev), col_bias=b3.to(cuda_dev), act=ops.ACT_RELU, out=out[:, 64:])
    ref = torch.cat([(a.float() @ w1.float().t() + b1).relu(), (a3.float() @ w3.float().t() + b3).relu()], 1)
    got = out.float().cpu()
    assert float((got - ref).abs().max()) < 2e-2 and float(got.min()) == 0.0
    assert float(((got == 0) != (ref == 0)).float().mean()) < 1e-3


# ---- networks -----------------------------------------------------------------------------------------------------------------
def _hf_clip(vcfg, tcfg, params, dev):
    from transformers import CLIPConfig, CLIPModel
    hc = CLIPConfig(text_config=dict(vocab_size=tcfg.vocab_size, hidden_size=tcfg.hidden_size, intermediate_size=tcfg.intermediate_size,
                                     num_hidden_layers=tcfg.num_layers, num_attention_heads=tcfg.num_heads, max_position_embeddings=tcfg.max_positions,
                                     hidden_act=tcfg.hidden_act, layer_norm_eps=tcfg.layer_norm_eps, eos_token_id=2, bos_token_id=0, pad_token_id=1),
                    vision_config=dict(hidden_size=vcfg.hidden_size, intermediate_size=vcfg.intermediate_size, num_hidden_layers=vcfg.num_layers,
                                       num_attention_heads=vcfg.num_heads, image_size=vcfg.image_size, patch_size=vcfg.patch_size,
                                       hidden_act=vcfg.hidden_act, layer_norm_eps=vcfg.layer_norm_eps),
                    projection_dim=vcfg.projection_dim)
    m = CLIPModel(hc).eval()
    res = m.load_state_dict(params, strict=False)
    assert not res.unexpected_keys and all(("position_ids" in k or k == "logit_scale") for k in res.missing_keys), res
    return m.to(dev)


@pytest.mark.parametrize("which", ["tiny", "vit-b16"])
def test_clip_image_tower_matches_transformers(cuda_dev, which):
    from fast_image_editing_with_generative_models_b200 import vit
    from fast_image_editing_with_generative_models_b200.metrics import clip_b16_text_config
    from fast_image_editing_with_generative_models_b200.text_encoder import make_clip_params, tiny_clip_config
    vcfg = vit.tiny_vit_config("clip") if which == "tiny" else vit.clip_b16_vision_config()
    tcfg = tiny_clip_config(True) if which == "tiny" else clip_b16_text_config()
    params = dict(vit.make_vit_params(vcfg)); params.update(make_clip_params(tcfg))
    hf = _hf_clip(vcfg, tcfg, params, cuda_dev)
    s = vcfg.image_size
    imgs = np.stack([_img(20, s, s), _img(21, s, s, False)])
    px = torch.cat([MO.clip_preprocess(i, s) for i in imgs]).to(cuda_dev)
    with torch.no_grad():
        ref = hf.visual_projection(hf.vision_model(pixel_values=px).pooler_output).float().cpu()
    got = vit.VisionTransformer(params, vcfg, cuda_dev).embed(_dev(imgs, cuda_dev), vit.CLIP_MEAN, vit.CLIP_STD).float().cpu()
    err, scale = float((got - ref).abs().max()), float(ref.abs().max())
    cos = F.cosine_similarity(got, ref, dim=1)
    print(f"\n[clip {which}] image_embeds max-abs {err:.4g} (ref absmax {scale:.3g}); cosine to the reference {cos.tolist()}")
    assert err <= 2e-2 * max(1.0, scale)                               # fp16 activations (the text-tower tolerance of test_gpu_text_encoder.py)
    assert float(cos.min()) > 0.9995


def test_clip_score_matches_oracle(cuda_dev):
    """MetricsCalculator.calculate_clip_score on a PIL image of arbitrary size against torchmetrics' formula on transformers' CLIPModel."""
    import warnings
    from PIL import Image
    from fast_image_editing_with_generative_models_b200 import vit
    from fast_image_editing_with_generative_models_b200.metrics import MetricsCalculator, clip_b16_text_config
    from fast_image_editing_with_generative_models_b200.text_encoder import make_clip_params, pseudo_token_ids
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        calc = MetricsCalculator(device="cuda")
    assert any("RANDOM weights" in str(x.message) for x in w) and len(calc.synthetic) == 3
    vcfg, tcfg = vit.clip_b16_vision_config(), clip_b16_text_config()
    params = dict(vit.make_vit_params(vcfg)); params.update(make_clip_params(tcfg))
    hf = _hf_clip(vcfg, tcfg, params, cuda_dev)
    for seed, (h, w_), text in ((30, (512, 512), "a photo of a red bicycle"), (31, (300, 400), "a watercolor painting of a fox in the snow"), (32, (224, 224), "")):
        img = _img(seed, h, w_)
        ids = pseudo_token_ids(text, tcfg.vocab_size).unsqueeze(0)
        ref = MO.clip_cosine(hf, img, ids)
        got = calc.clip_cosine(Image.fromarray(img), text)
        print(f"\n[clip score {h}x{w_}] cosine {got:.5f} vs oracle {ref:.5f}")
        assert abs(got - ref) < 5e-3                                    # i.e. 0.5 of a CLIPScore point (scores are 0..100)
        assert calc.calculate_clip_score(Image.fromarray(img), text) == max(100.0 * got, 0.0)


def test_dino_distance_matches_oracle(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import vit
    from fast_image_editing_with_generative_models_b200.metrics import DinoDistanceMetric
    for cfg, layer, size in ((vit.tiny_vit_config("dino", image_size=64, patch_size=8), 2, 64), (vit.dino_vitb8_config(), 11, 224)):
        params = vit.make_vit_params(cfg)
        metric = DinoDistanceMetric(device="cuda", resize_to=size, layer=layer, params=params, config=cfg)
        oracle = MO.DinoViT(params, cfg.patch_size, cfg.num_heads).to(cuda_dev)
        src = _img(40, 512, 512)
        # the similarity map itself: fp16 keys after `layer` blocks against the fp32 oracle
        (sim,), t = metric._self_similarity(_dev(src[None], cuda_dev))
        ref_sim = MO.dino_keys_self_similarity(oracle, MO.dino_preprocess(src, size).to(cuda_dev), layer)[0].cpu()
        e = sim[:, :t].cpu() - ref_sim
        rms, worst = float(e.pow(2).mean().sqrt()), float(e.abs().max())
        print(f"\n[dino {cfg.name}] {t} tokens; self-similarity error rms {rms:.3g}, max {worst:.3g}")
        assert t == cfg.num_patches + 1 and rms < 2e-3 and worst < 2e-2
        for edited in (_noisy(src, 41, 25), _img(42, 512, 512)):
            ref = MO.dino_distance(oracle, src, edited, layer=layer, resize_to=size)
            got = metric.calculate_distance(src, edited)
            print(f"[dino {cfg.name}] distance {got:.6g} vs oracle {ref:.6g}")
            # d = mean((delta + e)^2) with |e|_rms <= 2 * 2e-3 (two maps): |d - mean(delta^2)| <= 2 sqrt(d_ref) |e|_rms + |e|_rms^2
            assert abs(got - ref) <= 2 * math.sqrt(ref) * 4e-3 + 1.6e-5
        assert metric.calculate_distance(src, src) == 0.0


def test_lpips_matches_oracle(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import lpips as L
    params = L.make_lpips_params()
    net = L.LPIPSSqueeze(params, cuda_dev)
    feats = MO.squeezenet_features(params).to(cuda_dev)
    lins = [params[f"lin{k}.model.1.weight"].reshape(-1) for k in range(7)]
    a = np.stack([_img(50), _img(51), _img(52, 96, 80)[:64, :64].repeat(8, 0).repeat(8, 1)])
    b = np.stack([_noisy(a[0], 53, 20), _img(54), a[2]])
    got = net.distance(_dev(a, cuda_dev), _dev(b, cuda_dev)).cpu().numpy()
    for i in range(3):
        ref = MO.lpips_squeeze(feats, lins, a[i], b[i])
        print(f"\n[lpips {i}] {got[i]:.6g} vs oracle {ref:.6g}")
        assert abs(got[i] - ref) <= 2e-2 * ref + 1e-7                   # fp16 activations through up to 8 fire modules
    assert got[2] == 0.0
    # the tapped activations themselves
    taps = net.features(_dev(a[:1], cuda_dev))
    x = ((MO._to_float_nchw(a[0]).to(cuda_dev) * 2 - 1) - torch.tensor(MO.LPIPS_SHIFT, device=cuda_dev).view(1, 3, 1, 1)) / torch.tensor(MO.LPIPS_SCALE, device=cuda_dev).view(1, 3, 1, 1)
    k = 0
    with torch.no_grad():
        for i, layer in enumerate(feats):
            x = layer(x)
            if i in MO.LPIPS_TAPS:
                ref = x.permute(0, 2, 3, 1)
                assert taps[k].shape == ref.shape, (i, taps[k].shape, ref.shape)
                err = float((taps[k].float() - ref).abs().max())
                assert err <= 2e-2 * max(1.0, float(ref.abs().max())), (i, err)
                k += 1
    assert k == 7 and [t.shape[-1] for t in taps] == list(L.TAP_CHANNELS) and [t.shape[1] for t in taps] == [255, 127, 63, 31, 31, 31, 31]


def test_metrics_calculator_all_metrics_on_pil_images(cuda_dev):
    """The reference's call: ``calculate_all_metrics(source PIL, edited PIL, prompt)`` on 1024^2 edits of 512^2 sources."""
    import warnings
    from PIL import Image
    from fast_image_editing_with_generative_models_b200.metrics import MetricsCalculator
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        calc = MetricsCalculator(device="cuda")
    src = _img(60, 512, 512)
    big = np.asarray(Image.fromarray(_noisy(src, 61, 30)).resize((1024, 1024), Image.LANCZOS))
    m = calc.calculate_all_metrics(Image.fromarray(src), Image.fromarray(big), "a photo")
    assert set(m) == {"ssim", "lpips", "clip_score", "psnr", "mse", "dino_distance"} and all(isinstance(v, float) for v in m.values())
    small = np.asarray(Image.fromarray(big).resize((512, 512), Image.LANCZOS))                  # what the reference feeds its metrics
    assert abs(m["ssim"] - MO.ssim(src, small, dtype=torch.float64)) < 1e-9
    assert abs(m["mse"] - MO.mse(src, small)) < 1e-15 and abs(m["psnr"] - MO.psnr(src, small)) < 1e-9
    assert m["lpips"] > 0 and m["dino_distance"] > 0 and m["clip_score"] >= 0 and math.isfinite(m["lpips"])
    same = calc.calculate_all_metrics(Image.fromarray(src), Image.fromarray(src), "a photo")
    assert abs(same["ssim"] - 1.0) < 1e-12 and same["mse"] == 0.0 and same["psnr"] == float("inf") and same["lpips"] == 0.0 and same["dino_distance"] == 0.0
    # tensors in the layouts the reference's helpers accept (CHW float in [0, 1])
    t = torch.from_numpy(src).permute(2, 0, 1).float() / 255.0
    assert calc.calculate_mse(t, Image.fromarray(src)) == 0.0
    calc.clear_memory()


def test_metrics_from_checkpoint_files_written_by_the_real_libraries(cuda_dev, tmp_path):
    """``MetricsCalculator(checkpoints=...)`` on files produced by transformers (``CLIPModel.save_pretrained``), torchvision
    (``squeezenet1_1().state_dict()``) and a DINO-named ``.pth``; the scores must match the modules that wrote the files."""
    import warnings
    import torchvision
    from transformers import CLIPConfig, CLIPModel
    from fast_image_editing_with_generative_models_b200 import lpips as L
    from fast_image_editing_with_generative_models_b200 import vit
    from fast_image_editing_with_generative_models_b200.metrics import DinoDistanceMetric, MetricsCalculator
    from fast_image_editing_with_generative_models_b200.text_encoder import pseudo_token_ids
    torch.manual_seed(7)
    hc = CLIPConfig(text_config=dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, max_position_embeddings=77,
                                     eos_token_id=2, bos_token_id=0, pad_token_id=1),
                    vision_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=64, patch_size=16),
                    projection_dim=64)
    hf = CLIPModel(hc).eval()
    hf.save_pretrained(tmp_path / "clip")
    sq = torchvision.models.squeezenet1_1(weights=None).eval()
    torch.save(sq.state_dict(), tmp_path / "squeezenet1_1.pth")
    lins = {f"lin{k}.model.1.weight": torch.rand(1, c, 1, 1) * (2.0 / c) for k, c in enumerate(L.TAP_CHANNELS)}
    torch.save(lins, tmp_path / "lpips.pth")
    dcfg = vit.tiny_vit_config("dino", image_size=64, patch_size=8)
    dparams = vit.make_vit_params(dcfg)
    torch.save(dparams, tmp_path / "dino.pth")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        calc = MetricsCalculator(device="cuda", checkpoints={"clip": str(tmp_path / "clip"), "squeezenet": str(tmp_path / "squeezenet1_1.pth"), "lpips": str(tmp_path / "lpips.pth")})
    assert calc.synthetic == ["DINO ViT-B/8"] and any("pseudo token ids" in str(x.message) for x in w)       # (no vocab.json next to the model)
    a, b = _img(70), _noisy(_img(70), 71, 25)
    text = "a photo of a cat"
    got = calc.clip_cosine(a, text)
    ref = MO.clip_cosine(hf.to(cuda_dev), a, pseudo_token_ids(text, 1000).unsqueeze(0))
    print(f"\n[files] clip cosine {got:.5f} vs transformers {ref:.5f}")
    assert abs(got - ref) < 5e-3
    got = calc.calculate_lpips(a, b)
    ref = MO.lpips_squeeze(sq.features.to(cuda_dev), [lins[f"lin{k}.model.1.weight"].reshape(-1) for k in range(7)], a, b)
    print(f"[files] lpips {got:.6g} vs torchvision backbone {ref:.6g}")
    assert abs(got - ref) <= 2e-2 * ref
    metric = DinoDistanceMetric(device="cuda", checkpoint=str(tmp_path / "dino.pth"), resize_to=64, layer=2)      # architecture read off the tensor shapes
    assert (metric.model.cfg.patch_size, metric.model.cfg.image_size, metric.model.cfg.num_layers, metric.model.cfg.num_heads) == (8, 64, 3, 2)
    ref = MO.dino_distance(MO.DinoViT(dparams, 8, 2).to(cuda_dev), a, b, layer=2, resize_to=64)
    got = metric.calculate_distance(a, b)
    print(f"[files] dino distance {got:.6g} vs oracle {ref:.6g}")
    assert abs(got - ref) <= 2 * math.sqrt(ref) * 4e-3 + 1.6e-5

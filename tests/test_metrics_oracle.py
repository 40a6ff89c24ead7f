"""CPU: the evaluation-metric oracle (oracle/metrics_oracle.py) against what this image can pin it with — an independent
float64 SSIM, exact integer MSE, transformers' ViT (the HF port of DINO) for the DINO restatement, Pillow / torch for the two
resamplers' host tables, transformers' CLIP modules for the parameter naming — and the host logic of ``evaluate.py``."""
import csv
import json
import math

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

from fast_image_editing_with_generative_models_b200 import lpips as lpips_mod
from fast_image_editing_with_generative_models_b200 import vit
from fast_image_editing_with_generative_models_b200.resize import aa_bilinear_tables, pillow_tables


def _img(seed, h=96, w=80, smooth=True):
    rng = np.random.RandomState(seed)
    x = rng.rand(h, w, 3)
    if smooth:                                   # natural-image-like correlation so local variances are not all alike
        for _ in range(3):
            x = (x + np.roll(x, 1, 0) + np.roll(x, 1, 1) + np.roll(x, -1, 0)) / 4
        x = (x - x.min()) / (x.max() - x.min())
    return (x * 255).astype(np.uint8)


def test_ssim_restatement_equals_independent_valid_window_evaluation():
    a = _img(1)
    for b in (_img(2), np.clip(a.astype(int) + np.random.RandomState(3).randint(-12, 13, a.shape), 0, 255).astype(np.uint8), a[:, ::-1].copy()):
        s32, s64 = MO.ssim(a, b), MO.ssim_valid_window_f64(a, b)
        assert abs(s32 - s64) < 2e-5, (s32, s64)
        assert abs(MO.ssim(a, b, dtype=torch.float64) - s64) < 1e-7      # (the pixels are divided by 255 in float32 first, as the reference does)
    assert abs(MO.ssim(a, a) - 1.0) < 1e-6
    flat = np.full((40, 40, 3), 77, np.uint8)
    assert abs(MO.ssim(flat, flat, dtype=torch.float64) - 1.0) < 1e-9      # zero variance: the c2 terms keep the ratio at 1 ...
    assert abs(MO.ssim(flat, flat) - 1.0) < 1e-3          # ... but float32 moments (torchmetrics' arithmetic) leave ~1e-4 of cancellation noise


def test_mse_psnr_definitions():
    a = _img(4, smooth=False)
    b = np.clip(a.astype(int) + 1, 0, 255).astype(np.uint8)
    changed = (a != b).mean()
    assert abs(MO.mse(a, b) - changed / 255.0 ** 2) < 1e-15
    assert abs(MO.psnr(a, b) - 10 * math.log10(255.0 ** 2 / changed)) < 1e-9
    assert MO.psnr(a, a) == float("inf") and MO.mse(a, a) == 0.0
    ref = float(((torch.from_numpy(a).float() / 255 - torch.from_numpy(b).float() / 255) ** 2).mean())      # the reference's float32 route
    assert abs(MO.mse(a, b) - ref) < 1e-9


def _hf_vit_from_dino(params, cfg):
    from transformers import ViTConfig, ViTModel
    hc = ViTConfig(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, intermediate_size=cfg.intermediate_size,
                   hidden_act="gelu", layer_norm_eps=cfg.layer_norm_eps, image_size=cfg.image_size, patch_size=cfg.patch_size, qkv_bias=True,
                   hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    m = ViTModel(hc, add_pooling_layer=False).eval()
    c = cfg.hidden_size
    sd = {"embeddings.cls_token": params["cls_token"], "embeddings.position_embeddings": params["pos_embed"],
          "embeddings.patch_embeddings.projection.weight": params["patch_embed.proj.weight"],
          "embeddings.patch_embeddings.projection.bias": params["patch_embed.proj.bias"],
          "layernorm.weight": params["norm.weight"], "layernorm.bias": params["norm.bias"]}
    for i in range(cfg.num_layers):
        s, d = f"blocks.{i}.", f"encoder.layer.{i}."
        for j, n in enumerate(("query", "key", "value")):
            sd[d + f"attention.attention.{n}.weight"] = params[s + "attn.qkv.weight"][j * c:(j + 1) * c]
            sd[d + f"attention.attention.{n}.bias"] = params[s + "attn.qkv.bias"][j * c:(j + 1) * c]
        for a, b in (("attention.output.dense", "attn.proj"), ("layernorm_before", "norm1"), ("layernorm_after", "norm2"),
                     ("intermediate.dense", "mlp.fc1"), ("output.dense", "mlp.fc2")):
            sd[d + a + ".weight"], sd[d + a + ".bias"] = params[s + b + ".weight"], params[s + b + ".bias"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if "pooler" not in k], (missing, unexpected)
    return m


def test_dino_restatement_equals_transformers_vit():
    cfg = vit.tiny_vit_config("dino", image_size=32, patch_size=8)
    params = vit.make_vit_params(cfg)
    oracle = MO.DinoViT(params, cfg.patch_size, cfg.num_heads)
    hf = _hf_vit_from_dino(params, cfg)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    grabbed = {}
    hook = hf.encoder.layer[2].attention.attention.key.register_forward_hook(lambda m, i, o: grabbed.setdefault("k", o.detach()))
    with torch.no_grad():
        out, qkvs = oracle(x, capture_qkv=True)
        ref = hf(pixel_values=x).last_hidden_state
    hook.remove()
    assert float((out - ref).abs().max()) < 2e-5
    c = cfg.hidden_size
    assert float((qkvs[2][..., c:2 * c] - grabbed["k"]).abs().max()) < 2e-5        # the key third of the fused projection == HF's key Linear


def test_dino_self_similarity_shape_and_symmetry():
    cfg = vit.tiny_vit_config("dino", image_size=32, patch_size=8)
    m = MO.DinoViT(vit.make_vit_params(cfg), cfg.patch_size, cfg.num_heads)
    s = MO.dino_keys_self_similarity(m, MO.dino_preprocess(_img(5, 64, 64), 32), layer=2)
    assert s.shape == (1, 17, 17)
    assert float((s - s.transpose(1, 2)).abs().max()) < 1e-6 and float((torch.diagonal(s[0]) - 1).abs().max()) < 1e-5
    assert MO.dino_distance(m, _img(5, 64, 64), _img(5, 64, 64), layer=2, resize_to=32) == 0.0
    assert MO.dino_distance(m, _img(5, 64, 64), _img(6, 64, 64), layer=2, resize_to=32) > 0.0


def test_vit_clip_parameter_names_are_transformers():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = vit.tiny_vit_config("clip")
    hc = CLIPVisionConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                          image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act=cfg.hidden_act, layer_norm_eps=cfg.layer_norm_eps, projection_dim=cfg.projection_dim)
    m = CLIPVisionModelWithProjection(hc)
    want = {k: tuple(v.shape) for k, v in m.state_dict().items() if "position_ids" not in k}
    got = {k: tuple(v.shape) for k, v in vit.make_vit_params(cfg).items()}
    assert got == want
    full = vit.make_vit_params(vit.ViTConfig(num_layers=1))             # ViT-B/16 widths: 197 positions, 768-wide patch kernel
    assert full["vision_model.embeddings.position_embedding.weight"].shape == (197, 768)
    assert full["vision_model.embeddings.patch_embedding.weight"].shape == (768, 3, 16, 16) and full["visual_projection.weight"].shape == (512, 768)
    assert vit.make_vit_params(vit.ViTConfig(style="dino", patch_size=8, num_layers=1, projection_dim=None))["pos_embed"].shape == (1, 785, 768)


def _emulate_pillow(img, oh, ow, filt):
    def one_pass(a, out_size, axis):
        b, k, _ = pillow_tables(a.shape[axis], out_size, filt)
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((out_size,) + a.shape[1:], dtype=np.uint8)
        for i in range(out_size):
            x0, cnt = b[i].tolist()
            out[i] = np.clip(((1 << 21) + np.tensordot(k[i, :cnt].numpy().astype(np.int64), a[x0:x0 + cnt], axes=(0, 0))) >> 22, 0, 255)
        return np.moveaxis(out, 0, axis)
    t = one_pass(img, ow, 1) if ow != img.shape[1] else img
    return one_pass(t, oh, 0) if oh != img.shape[0] else t


@pytest.mark.parametrize("h,w,oh,ow", [(512, 512, 224, 224), (300, 200, 336, 224), (100, 120, 224, 268), (224, 224, 224, 224)])
def test_bicubic_tables_reproduce_pillow(h, w, oh, ow):
    from PIL import Image
    img = _img(h + w, h, w, smooth=False)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(_emulate_pillow(img, oh, ow, "bicubic"), ref)


@pytest.mark.parametrize("n_in,n_out", [(512, 224), (1024, 224), (300, 224), (100, 224), (224, 224)])
def test_antialias_tables_reproduce_torch(n_in, n_out):
    b, k, ks = aa_bilinear_tables(n_in, n_out)
    x = torch.rand(1, 3, n_in, 5, generator=torch.Generator().manual_seed(n_in))
    ref = torch.nn.functional.interpolate(x, size=(n_out, 5), mode="bilinear", antialias=True, align_corners=False)
    out = torch.zeros_like(ref)
    for i in range(n_out):
        x0, cnt = b[i].tolist()
        out[:, :, i] = (x[:, :, x0:x0 + cnt] * k[i, :cnt].view(1, 1, -1, 1)).sum(2)
    assert float((out - ref).abs().max()) < 1e-6
    assert k.shape == (n_out, ks) and float((k.sum(1) - 1).abs().max()) < 1e-6


def test_clip_preprocess_is_pillow_bicubic_centre_crop():
    img = _img(9, 300, 400, smooth=False)
    px = MO.clip_preprocess(img, 224)
    assert px.shape == (1, 3, 224, 224)
    from PIL import Image
    ref = np.asarray(Image.fromarray(img).resize((298, 224), Image.BICUBIC))[:, 37:37 + 224].astype(np.float32) / 255
    ref = (ref - np.asarray(MO.CLIP_MEAN, np.float32)) / np.asarray(MO.CLIP_STD, np.float32)
    assert np.array_equal(px[0].permute(1, 2, 0).numpy(), ref)


def test_lpips_oracle_backbone_is_torchvision_and_head_properties():
    params = lpips_mod.make_lpips_params()
    feats = MO.squeezenet_features(params)
    import torchvision
    ref = torchvision.models.squeezenet1_1(weights=None)
    assert [type(a) for a in feats] == [type(a) for a in ref.features]
    assert {k for k in params if k.startswith("features.")} == {"features." + k for k in ref.features.state_dict()}
    lins = [params[f"lin{k}.model.1.weight"].reshape(-1) for k in range(7)]
    assert [l.numel() for l in lins] == list(lpips_mod.TAP_CHANNELS)
    a, b = _img(10, 64, 64), _img(11, 64, 64)
    assert MO.lpips_squeeze(feats, lins, a, a) == 0.0
    d_ab, d_ba = MO.lpips_squeeze(feats, lins, a, b), MO.lpips_squeeze(feats, lins, b, a)
    assert d_ab > 0 and abs(d_ab - d_ba) < 1e-7
    near = np.clip(a.astype(int) + np.random.RandomState(0).randint(-2, 3, a.shape), 0, 255).astype(np.uint8)
    assert MO.lpips_squeeze(feats, lins, a, near) < d_ab


class _FakeCalculator:
    def __init__(self):
        self.calls = []

    def to_metric_size(self, img):
        return img

    def calculate_all_metrics(self, source_img, edited_img, prompt):
        self.calls.append(prompt)
        v = float(len(self.calls))
        return {"ssim": 0.5 + v / 100, "lpips": v / 10, "clip_score": 20 + v, "psnr": 10 + v, "mse": v / 1000, "dino_distance": v / 50}


def test_evaluate_cli_writes_the_reference_csv_and_summary(tmp_path):
    from PIL import Image
    import evaluate
    src, out = tmp_path / "src", tmp_path / "outputs" / "sdxl_fp16"
    (src / "0_random").mkdir(parents=True)
    (out / "0_random").mkdir(parents=True)
    mapping = {}
    for i in range(5):
        rel = f"0_random/{i:03d}.jpg"
        Image.fromarray(_img(i, 32, 32)).save(src / rel)
        if i != 3:                                                         # one output is missing -> skipped, not fatal
            Image.fromarray(_img(i + 50, 32, 32)).save(out / rel)
        mapping[f"{i:012d}"] = {"image_path": rel, "editing_prompt": f"prompt {i}", "editing_type_id": str(i % 2)}
    (tmp_path / "map.json").write_text(json.dumps(mapping))
    calc = _FakeCalculator()
    summary = evaluate.main(["--mapping_file", str(tmp_path / "map.json"), "--source_dir", str(src), "--outputs_dir", str(out),
                             "--results_file", str(tmp_path / "res" / "metrics.csv"), "--summary_file", str(tmp_path / "res" / "summary.json")], calculator=calc)
    rows = list(csv.DictReader(open(tmp_path / "res" / "metrics.csv")))
    assert list(rows[0]) == ["image_id", "image_path", "editing_type_id", "editing_prompt", "ssim", "lpips", "clip_score", "psnr", "mse", "dino_distance"]
    assert len(rows) == 4 and calc.calls == ["prompt 0", "prompt 1", "prompt 2", "prompt 4"]
    on_disk = json.load(open(tmp_path / "res" / "summary.json"))
    assert on_disk == summary and summary["total_images"] == 4
    assert set(summary["overall"]["ssim"]) == {"mean", "std", "median"} and set(summary["by_category"]) == {"0", "1"}
    assert summary["by_category"]["0"]["count"] == 3 and set(summary["by_category"]["1"]["psnr"]) == {"mean", "std"}
    args = evaluate.parse_args(["--outputs_dir", "outputs/batch/edited/ssd-1b_fp16/"])
    assert args.results_file == "results/ssd-1b_fp16/metrics.csv" and args.summary_file == "results/ssd-1b_fp16/summary.json"
    assert evaluate.parse_args(["--outputs_dir", "somewhere"]).results_file == "results/metrics.csv"


def test_metrics_calculator_refuses_to_run_without_cuda():
    from src.metrics import DinoDistanceMetric, MetricsCalculator          # the reference's import path
    assert DinoDistanceMetric is not None
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MetricsCalculator(device="cuda")


def test_metric_checkpoint_formats_written_by_the_real_libraries(tmp_path, monkeypatch):
    """The on-disk formats MetricsCalculator reads, produced here by the libraries that publish them: ``CLIPModel.save_pretrained``
    (transformers), ``torch.save(squeezenet1_1().state_dict())`` (torchvision), a DINO-named ``.pth`` — keys and shapes must be exactly
    what the towers consume (constructed on the CPU: no kernel runs)."""
    import torchvision
    from transformers import CLIPConfig, CLIPModel
    from fast_image_editing_with_generative_models_b200 import metrics as M
    from fast_image_editing_with_generative_models_b200.text_encoder import CLIPTextEncoder
    hc = CLIPConfig(text_config=dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=77),
                    vision_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=64, patch_size=16),
                    projection_dim=64)
    CLIPModel(hc).save_pretrained(tmp_path / "clip-vit-base-patch16")
    vcfg, tcfg, sd = M.load_clip_model_dir(str(tmp_path / "clip-vit-base-patch16"))
    assert (vcfg.hidden_size, vcfg.num_layers, vcfg.image_size, vcfg.patch_size, vcfg.projection_dim) == (128, 2, 64, 16, 64)
    assert (tcfg.hidden_size, tcfg.num_layers, tcfg.num_heads, tcfg.vocab_size, tcfg.projection_dim, tcfg.hidden_act) == (128, 2, 2, 1000, 64, "quick_gelu")
    tower = vit.VisionTransformer(sd, vcfg, "cpu")
    assert tower.w_patch.shape == (128, 16 * 16 * 3) and tower.pos.shape == (17, 128) and tower.proj.shape == (64, 128) and len(tower.layers) == 2
    text = CLIPTextEncoder(sd, tcfg, "cpu")
    assert text.proj.shape == (64, 128) and len(text.layers) == 2
    # torchvision's SqueezeNet + lpips' lin layers, DINO's hub checkpoint layout
    torch.save(torchvision.models.squeezenet1_1(weights=None).state_dict(), tmp_path / "squeezenet1_1.pth")
    lp = lpips_mod.make_lpips_params()
    torch.save({k: v for k, v in lp.items() if k.startswith("lin")}, tmp_path / "lpips_squeeze.pth")
    dcfg = vit.tiny_vit_config("dino", image_size=32, patch_size=8)
    torch.save(vit.make_vit_params(dcfg), tmp_path / "dino_vitbase8_pretrain.pth")
    monkeypatch.setenv("FIE_METRIC_CHECKPOINTS", str(tmp_path))
    found = M.metric_checkpoints_from_env()
    assert set(found) == {"clip", "dino", "squeezenet", "lpips"}
    p = {k: v.float() for k, v in torch.load(found["squeezenet"], weights_only=True).items()}
    p.update(torch.load(found["lpips"], weights_only=True))
    net = lpips_mod.LPIPSSqueeze(p, "cpu")                      # ignores torchvision's classifier.* entries, finds every features.* / lin* key
    assert net.w0.shape == (64, 64) and len(net.fires) == 8 and [l.numel() for l in net.lins] == list(lpips_mod.TAP_CHANNELS)
    assert vit.VisionTransformer(torch.load(found["dino"], weights_only=True), dcfg, "cpu").b_patch.shape == (128,)
    monkeypatch.delenv("FIE_METRIC_CHECKPOINTS")
    assert M.metric_checkpoints_from_env() == {}


def test_clip_cosine_restatement_equals_transformers_forward():
    """``CLIPModel.forward`` returns ``logits_per_image = exp(logit_scale) * cos(image_embeds, text_embeds)`` computed by transformers itself
    (projection, pooling and L2 normalisation included); the oracle's hand-assembled cosine must agree — and CLIPScore is 100 x that, floored."""
    from transformers import CLIPConfig, CLIPModel
    from fast_image_editing_with_generative_models_b200.text_encoder import pseudo_token_ids
    torch.manual_seed(3)
    hc = CLIPConfig(text_config=dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=77,
                                     eos_token_id=2, bos_token_id=0, pad_token_id=1),
                    vision_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=64, patch_size=16),
                    projection_dim=64)
    m = CLIPModel(hc).eval()
    img = _img(12, 200, 150)
    ids = pseudo_token_ids("a photo of a dog", 1000).unsqueeze(0)
    with torch.no_grad():
        out = m(input_ids=ids, pixel_values=MO.clip_preprocess(img, 64))
    ref = float(out.logits_per_image[0, 0] / m.logit_scale.exp())
    got = MO.clip_cosine(m, img, ids)
    assert abs(got - ref) < 1e-5, (got, ref)
    assert MO.clip_score(m, img, ids) == max(100.0 * got, 0.0)


def test_ssim_window_equals_scipy_gaussian_filter():
    """A third evaluation with a library Gaussian: ``scipy.ndimage.gaussian_filter(sigma=1.5, truncate=3.5)`` is the same 11-tap normalised
    window (radius int(3.5 * 1.5 + 0.5) = 5); away from the border, which torchmetrics crops, the SSIM map must coincide."""
    from scipy.ndimage import gaussian_filter
    a, b = _img(13), _img(14)
    x, y = (a.astype(np.float32) / 255.0).astype(np.float64), (b.astype(np.float32) / 255.0).astype(np.float64)
    f = lambda m: np.stack([gaussian_filter(m[..., c], sigma=1.5, truncate=3.5) for c in range(3)], -1)
    mx, my = f(x), f(y)
    vx, vy, cxy = np.maximum(f(x * x) - mx * mx, 0), np.maximum(f(y * y) - my * my, 0), f(x * y) - mx * my
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    full = ((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
    ref = float(full[5:-5, 5:-5].mean())
    assert abs(MO.ssim(a, b, dtype=torch.float64) - ref) < 1e-9

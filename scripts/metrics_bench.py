"""Timing of the evaluation metrics: ``MetricsCalculator.calculate_all_metrics`` per image pair on the GPU (host PIL images
in, Python floats out — the reference's call) with the per-metric split, next to the CPU restatement (oracle/metrics_oracle.py: torch
fp32 on all host cores, transformers' CLIP, torchvision's SqueezeNet) on the same pair.  Prints one JSON object.

    python scripts/metrics_bench.py [--pairs 20] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _img(seed, h=512, w=512):
    rng = np.random.RandomState(seed)
    x = rng.rand(h, w, 3)
    for _ in range(3):
        x = (x + np.roll(x, 1, 0) + np.roll(x, 1, 1) + np.roll(x, -1, 0)) / 4
    x = (x - x.min()) / (x.max() - x.min())
    return (x * 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from PIL import Image
    from fast_image_editing_with_generative_models_b200 import ops
    from fast_image_editing_with_generative_models_b200.metrics import MetricsCalculator
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        calc = MetricsCalculator(device="cuda")
    src = [Image.fromarray(_img(i)) for i in range(4)]
    edited = [Image.fromarray(_img(100 + i, 1024, 1024)) for i in range(4)]        # the editor's 1024^2 outputs; metrics run on 512^2 Lanczos copies
    prompt = "a watercolor painting of a fox in the snow"
    for i in range(3):
        calc.calculate_all_metrics(src[i % 4], edited[i % 4], prompt)
    torch.cuda.synchronize()
    launches0 = ops.LAUNCHES
    t0 = time.perf_counter()
    for i in range(args.pairs):
        calc.calculate_all_metrics(src[i % 4], edited[i % 4], prompt)
    torch.cuda.synchronize()
    per_pair = (time.perf_counter() - t0) / args.pairs
    launches = (ops.LAUNCHES - launches0) / args.pairs
    split = {}
    for name, fn in (("ssim", lambda: calc.calculate_ssim(src[0], edited[0])), ("lpips", lambda: calc.calculate_lpips(src[0], edited[0])),
                     ("clip_score", lambda: calc.calculate_clip_score(edited[0], prompt)), ("psnr", lambda: calc.calculate_psnr(src[0], edited[0])),
                     ("mse", lambda: calc.calculate_mse(src[0], edited[0])), ("dino_distance", lambda: calc.dino_metric.calculate_distance(src[0], edited[0]))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        split[name] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
    out = {"metric": "evaluated image pairs per second (all six metrics, PIL in -> floats out)", "value": round(1.0 / per_pair, 2), "ms_per_pair": round(per_pair * 1e3, 2),
           "gpu_launches_per_pair": launches, "ms_per_metric": split, "weights": "synthetic (seeded random init)", "pairs": args.pairs,
           "device": torch.cuda.get_device_name(0)}
    if not args.no_cpu:
        from oracle import metrics_oracle as MO
        from fast_image_editing_with_generative_models_b200 import lpips as L
        from fast_image_editing_with_generative_models_b200 import vit
        from fast_image_editing_with_generative_models_b200.metrics import clip_b16_text_config
        from fast_image_editing_with_generative_models_b200.text_encoder import make_clip_params, pseudo_token_ids
        from transformers import CLIPConfig, CLIPModel
        a = np.asarray(src[0])
        b = np.asarray(edited[0].resize((512, 512), Image.LANCZOS))
        cpu = {}

        def timed(name, fn, reps=2):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            cpu[name] = round((time.perf_counter() - t0) / reps * 1e3, 1)
        timed("resize_lanczos", lambda: edited[0].resize((512, 512), Image.LANCZOS))
        timed("ssim", lambda: MO.ssim(a, b))
        timed("mse+psnr", lambda: (MO.mse(a, b), MO.psnr(a, b)))
        lp = L.make_lpips_params()
        feats, lins = MO.squeezenet_features(lp), [lp[f"lin{k}.model.1.weight"].reshape(-1) for k in range(7)]
        timed("lpips", lambda: MO.lpips_squeeze(feats, lins, a, b))
        vcfg, tcfg = vit.clip_b16_vision_config(), clip_b16_text_config()
        params = dict(vit.make_vit_params(vcfg)); params.update(make_clip_params(tcfg))
        hc = CLIPConfig(text_config=dict(vocab_size=tcfg.vocab_size, hidden_size=tcfg.hidden_size, intermediate_size=tcfg.intermediate_size, num_hidden_layers=tcfg.num_layers,
                                         num_attention_heads=tcfg.num_heads, eos_token_id=2, bos_token_id=0, pad_token_id=1),
                        vision_config=dict(hidden_size=vcfg.hidden_size, intermediate_size=vcfg.intermediate_size, num_hidden_layers=vcfg.num_layers,
                                           num_attention_heads=vcfg.num_heads, image_size=224, patch_size=16), projection_dim=512)
        hf = CLIPModel(hc).eval()
        hf.load_state_dict(params, strict=False)
        ids = pseudo_token_ids(prompt, tcfg.vocab_size).unsqueeze(0)
        timed("clip_score", lambda: MO.clip_score(hf, b, ids))
        dcfg = vit.dino_vitb8_config()
        dino = MO.DinoViT(vit.make_vit_params(dcfg), 8, 12)
        timed("dino_distance", lambda: MO.dino_distance(dino, a, b), reps=1)
        cpu["total"] = round(sum(cpu.values()), 1)
        out["cpu_baseline"] = {"kind": "port", "cores": torch.get_num_threads(), "ms_per_metric": cpu, "value": round(1e3 / cpu["total"], 3),
                               "unit": "pairs/s", "sample": "one 512^2 / 1024^2 pair, torch fp32 on the host cores"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json configs[4] through the SHIPPED CLI: a PIE-Bench-shaped synthetic dataset on disk (700 JPEG sources of 512 x 512 in 10
categories 140/80/80/80/40/40/40/40/80/80 + mapping_file.json, as `/root/reference/run_batch.py:102-140` reads it) edited by
`run_batch.py` under torchrun on N GPUs, for SDXL and SSD-1B.  Per model it records the CLI's own counters (--summary_json: edit time =
max over ranks of the time spent inside FastEditor.edit_many, i.e. JPEG decode overlap, GPU Lanczos 512 -> 1024, the edit, the GPU JPEG
encode and D2H) and the outer wall clock (process start-up, synthetic weight generation and packing, graph capture, file writes).

    python scripts/sweep_cli_bench.py --gpus 8 --out gpurun_out/sweep700_n8.json [--models sdxl ssd-1b] [--num_images 700]
"""
import argparse
import json
import os
import subprocess
import sys
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CATEGORY_COUNTS = [140, 80, 80, 80, 40, 40, 40, 40, 80, 80]      # results/*/summary.json by_category of the reference


def _make(args):
    path, seed = args
    from PIL import Image
    from fast_image_editing_with_generative_models_b200.synthetic import synthetic_image
    Image.fromarray(synthetic_image(seed, 512, 512)).save(path, quality=90)
    return path


def make_dataset(root, n):
    os.makedirs(root, exist_ok=True)
    entries = [(cat, i) for cat, cnt in enumerate(CATEGORY_COUNTS) for i in range(cnt)][:n]
    mapping, jobs = {}, []
    for k, (cat, i) in enumerate(entries):
        rel = f"{cat}_category/{k:012d}.jpg"
        os.makedirs(os.path.join(root, "annotation_images", f"{cat}_category"), exist_ok=True)
        mapping[f"{k:012d}"] = {"image_path": rel, "original_prompt": "a photo", "editing_prompt": f"category {cat}: make image {i} a watercolour painting",
                                 "editing_type_id": str(cat)}
        jobs.append((os.path.join(root, "annotation_images", rel), k))
    with Pool(min(os.cpu_count() or 1, 32)) as p:
        p.map(_make, jobs, chunksize=8)
    mf = os.path.join(root, "mapping_file.json")
    with open(mf, "w") as f:
        json.dump(mapping, f)
    return mf, os.path.join(root, "annotation_images")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--models", nargs="+", default=["sdxl", "ssd-1b"])
    ap.add_argument("--num_images", type=int, default=sum(CATEGORY_COUNTS))
    ap.add_argument("--strength", type=float, default=0.5)
    ap.add_argument("--out", default="gpurun_out/sweep700.json")
    ap.add_argument("--data", default="/tmp/pie_synth")
    a = ap.parse_args()
    t0 = time.time()
    mf, src = make_dataset(a.data, a.num_images)
    t_data = time.time() - t0
    results = {"config": f"PIE-Bench-shaped synthetic sweep through run_batch.py: {a.num_images} JPEG sources of 512x512 -> 1024x1024 edits, fp16, "
                         f"4 LCM steps @ strength {a.strength}, CFG 1.5, micro-batch 8 per GPU, GPU JPEG encode, {a.gpus} GPU(s) of one box",
               "n_gpus": a.gpus, "images": a.num_images, "dataset_generation_s": round(t_data, 1), "models": {}}
    for model in a.models:
        out_dir = os.path.join(a.data, f"outputs_{model}")
        summ = os.path.join(a.data, f"summary_{model}.json")
        cmd = [sys.executable]
        if a.gpus > 1:
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29533"]
        cmd += [os.path.join(ROOT, "run_batch.py"), "--mapping_file", mf, "--source_dir", src, "--output_dir", out_dir, "--model", model,
                "--strength", str(a.strength), "--seed", "42", "--no_cpu_offload", "--summary_json", summ]
        t1 = time.time()
        rc = subprocess.call(cmd, cwd=ROOT, stdout=open(os.path.join(a.data, f"log_{model}.txt"), "w"), stderr=subprocess.STDOUT)
        wall = time.time() - t1
        r = {"rc": rc, "wall_s_including_startup": round(wall, 1)}
        if os.path.exists(summ):
            r.update(json.load(open(summ)))
        n_files = sum(len(fs) for _, _, fs in os.walk(os.path.join(out_dir, "batch", "edited")))
        r["files_written"] = n_files
        results["models"][model] = r
        print(model, json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(results, f, indent=1)
    print(json.dumps(results))


if __name__ == "__main__":
    main()

"""Evaluate edited images against their PIE-Bench sources — same command line, CSV columns and summary JSON as the reference's
``evaluate.py`` (``/root/reference/evaluate.py:25-300``), computed by the B200 ``MetricsCalculator``.

    python evaluate.py --outputs_dir outputs/batch/edited/sdxl_fp16 [--results_file results/metrics.csv]

Image files are decoded on a host thread pool while the GPU scores the previous pair."""
import argparse
import csv
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from PIL import Image

from src.metrics import MetricsCalculator

METRICS = ("ssim", "lpips", "clip_score", "psnr", "mse", "dino_distance")
KNOWN_SUFFIXES = ("sdxl_fp32", "sdxl_fp16", "ssd-1b_fp32", "ssd-1b_fp16")


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Evaluate edited images")
    ap.add_argument("--mapping_file", type=str, default="data/PIE-Bench_v1/mapping_file.json", help="Path to PIE-Bench mapping file")
    ap.add_argument("--source_dir", type=str, default="data/PIE-Bench_v1/annotation_images", help="Directory containing source images")
    ap.add_argument("--outputs_dir", type=str, required=True, help="Directory containing edited images (e.g., outputs/batch/edited/sdxl_fp32)")
    ap.add_argument("--results_file", type=str, default=None, help="Output CSV file for metrics (auto-detected from outputs_dir if not specified)")
    ap.add_argument("--summary_file", type=str, default=None, help="Output JSON file for summary statistics (auto-detected from outputs_dir if not specified)")
    ap.add_argument("--device", type=str, default="cuda", help="Device to use for metrics computation")
    ap.add_argument("--io_threads", type=int, default=4, help="host threads decoding image files ahead of the GPU")
    args = ap.parse_args(argv)
    tail = os.path.basename(args.outputs_dir.rstrip("/"))
    sub = f"results/{tail}" if tail in KNOWN_SUFFIXES else "results"
    args.results_file = args.results_file or f"{sub}/metrics.csv"
    args.summary_file = args.summary_file or f"{sub}/summary.json"
    return args


def _stats(values, median=False):
    out = {"mean": float(np.mean(values)), "std": float(np.std(values))}
    if median:
        out["median"] = float(np.median(values))
    return out


def summarise(rows):
    by_cat = {}
    for r in rows:
        by_cat.setdefault(r["editing_type_id"], []).append(r)
    summary = {"total_images": len(rows), "overall": {m: _stats([r[m] for r in rows], median=True) for m in METRICS}, "by_category": {}}
    for cat, rs in by_cat.items():
        summary["by_category"][cat] = {"count": len(rs), **{m: _stats([r[m] for r in rs]) for m in METRICS}}
    return summary


def _print_block(indent, d):
    print(f"{indent}SSIM:       {d['ssim']['mean']:.4f} ± {d['ssim']['std']:.4f}")
    print(f"{indent}LPIPS:      {d['lpips']['mean']:.4f} ± {d['lpips']['std']:.4f}")
    print(f"{indent}PSNR:       {d['psnr']['mean']:.2f} ± {d['psnr']['std']:.2f} dB")
    print(f"{indent}MSE:        {d['mse']['mean']:.6f} ± {d['mse']['std']:.6f}")
    print(f"{indent}CLIP Score: {d['clip_score']['mean']:.2f} ± {d['clip_score']['std']:.2f}")
    print(f"{indent}DINO Dist.: {d['dino_distance']['mean']:.4f} ± {d['dino_distance']['std']:.4f}")


def _load_pair(source_path, output_path):
    return Image.open(source_path).convert("RGB"), Image.open(output_path).convert("RGB")


def main(argv=None, calculator=None):
    args = parse_args(argv)
    for f in (args.results_file, args.summary_file):
        os.makedirs(os.path.dirname(f) or ".", exist_ok=True)
    print(f"\n[1/3] Loading mapping file from {args.mapping_file}")
    with open(args.mapping_file) as f:
        mapping = json.load(f)
    print(f"      Found {len(mapping)} entries in mapping file")
    print(f"\n[2/3] Scanning outputs directory: {args.outputs_dir}")
    if not os.path.isdir(args.outputs_dir):
        print(f"Error: Outputs directory not found or not a directory: {args.outputs_dir}")
        return None
    print(f"      Found {len(os.listdir(args.outputs_dir))} files in outputs directory")
    print("\n[3/3] Computing metrics...")
    calc = calculator or MetricsCalculator(device=args.device)

    todo, skipped = [], 0
    for image_id, entry in mapping.items():
        src, out = os.path.join(args.source_dir, entry["image_path"]), os.path.join(args.outputs_dir, entry["image_path"])
        if os.path.exists(src) and os.path.exists(out):
            todo.append((image_id, entry, src, out))
        else:
            skipped += 1
    rows = []
    with ThreadPoolExecutor(max(1, args.io_threads)) as pool:
        loads = [pool.submit(_load_pair, src, out) for _, _, src, out in todo]
        for (image_id, entry, _, _), fut in zip(todo, loads):
            try:
                source_img, edited_img = fut.result()
                prompt = entry.get("editing_prompt", "")
                # every metric is computed on 512^2 Lanczos copies (reference evaluate.py:127-141); the copies are made on the GPU,
                # bit-identical to Image.resize(..., Image.LANCZOS)
                m = calc.calculate_all_metrics(source_img=calc.to_metric_size(source_img), edited_img=calc.to_metric_size(edited_img), prompt=prompt)
                rows.append({"image_id": image_id, "image_path": entry["image_path"], "editing_type_id": entry.get("editing_type_id", "unknown"),
                             "editing_prompt": prompt, **{k: m[k] for k in METRICS}})
            except Exception as e:  # noqa: BLE001  (one bad file must not end the sweep; it is counted and reported)
                print(f"\n      Error processing {image_id}: {e}")
                skipped += 1
    print(f"\n      Processed: {len(rows)} images")
    print(f"      Skipped:   {skipped} images")
    if not rows:
        print("\n      No images were processed. Exiting.")
        return None

    print("\n[4/4] Saving results...")
    with open(args.results_file, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["image_id", "image_path", "editing_type_id", "editing_prompt", *METRICS])
        w.writeheader()
        w.writerows(rows)
    print(f"      Saved detailed metrics to: {args.results_file}")
    summary = summarise(rows)
    with open(args.summary_file, "w") as f:
        json.dump(summary, f, indent=2)
    print(f"      Saved summary statistics to: {args.summary_file}")

    bar = "=" * 60
    print(f"\n{bar}\nEVALUATION SUMMARY\n{bar}")
    print(f"\nTotal Images Evaluated: {len(rows)}\n\nOverall Metrics:")
    _print_block("  ", summary["overall"])
    print("\nMetrics by Category:")
    for cat in sorted(summary["by_category"], key=str):
        print(f"\n  Category {cat} ({summary['by_category'][cat]['count']} images):")
        _print_block("    ", summary["by_category"][cat])
    print(f"\n{bar}\n\nDone!")
    return summary


if __name__ == "__main__":
    main()

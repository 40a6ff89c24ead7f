"""Canny on the benchmark's synthetic 1024^2 images (batch 8 by default): CUDA-event time per call and the achieved fraction of the HBM
roofline (algorithmic bytes = 3 B/px RGB in + out_channels B/px out).  Target of `ncu -k regex:k_nms|k_seams|k_finalize`.
    python scripts/canny_probe.py [batch] [out_channels] [reps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_image_editing_with_generative_models_b200 import ops, synthetic as S   # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
imgs = torch.from_numpy(np.stack([S.synthetic_image(s, 1024, 1024) for s in range(batch)])).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    e = ops.canny(imgs, 100, 200, out_channels=ch)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()                                            # L2 flush between timed calls
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e = ops.canny(imgs, 100, 200, out_channels=ch); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
us = float(np.median(ts))
byts = imgs.numel() + e.numel()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6441.6
if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
print(json.dumps({"batch": batch, "out_channels": ch, "us_per_call": round(us, 1), "algorithmic_bytes": byts, "GB_per_s": round(byts / us / 1e3, 1),
                  "frac_of_hbm_peak": round(byts / us / 1e3 / peak, 4), "edge_density": round(float((e[..., 0] if ch == 3 else e).float().mean()) / 255, 4)}))

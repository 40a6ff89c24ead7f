"""Full-size timing probe with per-kernel-family breakdown (run on the GPU box)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fast_image_editing_with_generative_models_b200 import model_zoo, ops, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ssd-1b")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--tiny", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
t0 = time.time()
state = model_zoo.synthetic_state(a.model, tiny=a.tiny)
print("weights generated", time.time() - t0, flush=True)
eng = model_zoo.build_engine(state, dev)
del state
torch.cuda.synchronize()
print("engine packed", time.time() - t0, "mem GB", torch.cuda.memory_allocated() / 2**30, flush=True)
B, H = a.batch, a.size
imgs = torch.from_numpy(np.stack([S.synthetic_image(i, H, H) for i in range(B)])).to(dev)
ucfg = eng.unet.cfg
pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim)
noises = S.synthetic_noises(0, B, H // 8, H // 8)
for it in range(a.iters):
    torch.cuda.synchronize(); t = time.time()
    out = eng.edit_batch(imgs, pe, pl, noises)
    torch.cuda.synchronize(); dt = time.time() - t
    print(f"iter {it}: {dt*1e3:.1f} ms  ({B/dt:.2f} img/s)  peak mem GB {torch.cuda.max_memory_allocated()/2**30:.1f}", flush=True)
ops.PROFILE = []
l0 = ops.LAUNCHES
out = eng.edit_batch(imgs, pe, pl, noises)
summ = ops.profile_summary()
tags = ops.profile_summary(by_tag=True)
ops.PROFILE = None
print("launches per edit_batch", ops.LAUNCHES - l0)
tot = sum(d["ms"] for d in summ.values())
for fam, d in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    rate = d["work"] / (d["ms"] * 1e-3) if d["ms"] > 0 else 0
    print(f"{fam:14s} calls {d['calls']:5d}  {d['ms']:9.2f} ms  {100*d['ms']/tot:5.1f}%  " + (f"{rate/1e12:8.1f} TFLOP/s" if d["unit"] == "FLOP" else f"{rate/1e9:8.1f} GB/s"))
print("sum of profiled ms", tot)
print("--- top shapes ---")
for fam, d in sorted(tags.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    rate = d["work"] / (d["ms"] * 1e-3)
    print(f"{d['ms']:8.2f} ms x{d['calls']:4d}  " + (f"{rate/1e12:7.1f} TF/s" if d["unit"] == "FLOP" else f"{rate/1e9:7.1f} GB/s") + f"  {fam}")
print("img u8 mean/std", float(out.images.float().mean()), float(out.images.float().std()))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({k: {kk: vv for kk, vv in v.items()} for k, v in summ.items()}, open(f"gpurun_out/probe_{a.model}_b{B}.json", "w"), indent=1)

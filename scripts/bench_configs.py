"""The other BASELINE.json configurations (parity-test cases, not bench.py lines), timed once for DESIGN.md section 5:
   configs[1]  Canny + VAE encode/decode only, batch 32 x 1024x1024, fp16, 1 B200
   configs[2]  SSD-1B full edit, batch 1 (latency), eager vs CUDA-graph replay"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fast_image_editing_with_generative_models_b200 import model_zoo, ops, synthetic as S

dev = torch.device("cuda:0")
out = {}

def timed(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# ---- configs[2]: SSD-1B batch 1 latency
state = model_zoo.synthetic_state("ssd-1b")
eng = model_zoo.build_engine(state, dev)
ucfg = eng.unet.cfg
img = torch.from_numpy(np.stack([S.synthetic_image(0, 1024, 1024)])).to(dev)
pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim)
pe, pl = pe.to(dev), pl.to(dev)
nz = [n.to(dev, torch.float16) for n in S.synthetic_noises(0, 1, 128, 128)]
for graph in (False, True):
    ms = timed(lambda: eng.edit_batch(img, pe, pl, nz, strength=0.5, use_graph=graph), 10)
    out[f"ssd1b_b1_{'graph' if graph else 'eager'}_ms"] = ms
    print(f"configs[2] SSD-1B batch 1, strength 0.5: {ms:.1f} ms per edit ({'CUDA graph' if graph else 'eager'})", flush=True)

# ---- configs[1]: Canny + VAE encode/decode only, batch 32
B = 32
imgs = torch.from_numpy(np.stack([S.synthetic_image(i, 1024, 1024) for i in range(B)])).to(dev)
vae = eng.vae
def canny(): return ops.canny(imgs, 100, 200, out_channels=3)
def enc(): return vae.encode_moments(ops.preprocess_pad8(imgs, True))
lat = torch.randn((B, 128, 128, 4), device=dev).half()
def dec(): return ops.postprocess(vae.decode(lat))
for name, fn, it in (("canny", canny, 20), ("vae_encode", enc, 3), ("vae_decode", dec, 3)):
    ms = timed(fn, it)
    out[f"b32_{name}_ms"] = ms
    extra = f", {B * 1024 * 1024 * 6 / ms / 1e6:.0f} GB/s (3 B in + 3 B out per pixel)" if name == "canny" else ""
    print(f"configs[1] batch 32 {name}: {ms:.2f} ms ({B / ms * 1e3:.0f} images/s{extra})", flush=True)
tot = out["b32_canny_ms"] + out["b32_vae_encode_ms"] + out["b32_vae_decode_ms"]
print(f"configs[1] Canny + VAE encode + decode: {tot:.1f} ms per 32 images = {B / tot * 1e3:.1f} images/s  (peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB)")
out["b32_total_ms"] = tot
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)

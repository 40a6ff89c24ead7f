#!/usr/bin/env python
"""Host-side timeline of FastEditor.edit_many (where does the plugin path lose time against the device-resident engine call?).
    python scripts/e2e_probe.py [model] [batch] [n_batches]"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from fast_image_editing_with_generative_models_b200 import model_zoo, synthetic as S
from fast_image_editing_with_generative_models_b200.editor import FastEditor

model = sys.argv[1] if len(sys.argv) > 1 else "ssd-1b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 5
warnings.simplefilter("ignore")
state = model_zoo.synthetic_state(model)
ed = FastEditor(model_name=model, device="cuda", state=state, text_encoders=True, verbose=False)
pil = [Image.fromarray(S.synthetic_image(i, 1024, 1024)) for i in range(B)]
EDIT = dict(strength=0.5, num_inference_steps=4, guidance_scale=1.5, controlnet_conditioning_scale=0.5)


def run(nb, tag, **kw):
    imgs = pil * nb
    prompts = [f"{tag} prompt {time.time()} {j}" for j in range(len(imgs))]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ed.edit_many(imgs, prompts, seeds=list(range(len(imgs))), micro_batch=B, **EDIT, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{tag}: {nb} x {B} images in {dt * 1e3:.1f} ms = {dt * 1e3 / nb:.1f} ms per micro-batch", flush=True)
    return out


run(1, "warm-up")
run(1, "warm-up jpeg", output="jpeg")
for _ in range(2):
    run(1, "single")
    run(NB, "pipelined")
    run(NB, "pipelined jpeg", output="jpeg")
# phase timings of one micro-batch, each phase synchronised
ed._trace = []
run(NB, "traced")
tr = ed._trace
ed._trace = None
import collections
agg = collections.defaultdict(float)
for name, dt in tr:
    agg[name] += dt
for k, v in agg.items():
    print(f"   {k:28s} {v * 1e3 / NB:8.2f} ms per micro-batch (host time)")
# device-resident reference
eng = ed.pipe.engine
ucfg = eng.unet.cfg
pe, pl = S.synthetic_prompt(0, ucfg.cross_attention_dim, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim)
nz = [n.cuda().half() for n in S.synthetic_noises(0, B, 128, 128)]
imgs = torch.from_numpy(np.stack([np.array(p) for p in pil])).cuda()
for _ in range(2):
    eng.edit_batch(imgs, pe.cuda(), pl.cuda(), nz, strength=0.5)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(NB):
    eng.edit_batch(imgs, pe.cuda(), pl.cuda(), nz, strength=0.5)
torch.cuda.synchronize(); print(f"device-resident: {(time.perf_counter() - t0) * 1e3 / NB:.1f} ms per micro-batch")

"""BASELINE.json configs[4]: a PIE-Bench-shaped synthetic sweep (700 images in 10 categories: 140/80/80/80/40/40/40/40/80/80)
sharded over the GPUs of one box, as `run_batch.py` does it (rank r owns entries r, r+world, ...), in micro-batches of 8 images
per GPU; uint8 outputs are gathered over NCCL at the end.  Launch:  torchrun --nproc-per-node N scripts/sweep_bench.py [--model sdxl]
(or plain `python` for one GPU).  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fast_image_editing_with_generative_models_b200 import model_zoo, ops, sweep, synthetic as S

CATEGORY_COUNTS = [140, 80, 80, 80, 40, 40, 40, 40, 80, 80]      # results/*/summary.json by_category of the reference

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="sdxl", choices=["sdxl", "ssd-1b"])
ap.add_argument("--num_images", type=int, default=sum(CATEGORY_COUNTS))
ap.add_argument("--micro_batch", type=int, default=8)
a = ap.parse_args()
rank, world, local = sweep.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
entries = [(cat, i) for cat, n in enumerate(CATEGORY_COUNTS) for i in range(n)][: a.num_images]
mine = sweep.shard(list(range(len(entries))), rank, world)
eng = model_zoo.build_engine(model_zoo.synthetic_state(a.model), dev)
eng.use_graphs = True
ucfg = eng.unet.cfg
pooled_dim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
B = a.micro_batch

POOL = 32        # distinct synthetic source images per rank (generating 700 on the host would dominate the run); prompts and noise are per image
pool = torch.from_numpy(np.stack([S.synthetic_image(rank * POOL + j, 1024, 1024) for j in range(POOL)])).pin_memory()


def batch_inputs(ids):
    ids = list(ids) + [ids[-1]] * (B - len(ids))                        # pad the tail micro-batch (results dropped)
    imgs = pool[torch.tensor([i % POOL for i in ids])].pin_memory()
    pes, pls = zip(*[S.synthetic_prompt(zlib.crc32(f"cat{entries[i][0]}-img{entries[i][1]}".encode()) % (1 << 30), ucfg.cross_attention_dim, pooled_dim) for i in ids])
    nz = [torch.cat(t, 0) for t in zip(*[S.synthetic_noises(i, 1, 128, 128) for i in ids])]
    return imgs, torch.stack(pes), torch.stack(pls), nz                 # per-image prompts: [B,2,77,D], [B,2,P]

# warm-up (captures the graph), then the timed sweep
w = batch_inputs(mine[:B] if mine else [0])
eng.edit_batch(w[0].to(dev), w[1], w[2], w[3], strength=0.5)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
t0 = time.perf_counter()
outs = []
for k in range(0, len(mine), B):
    ids = mine[k:k + B]
    imgs, pe, pl, nz = batch_inputs(ids)
    out = eng.edit_batch(imgs.to(dev, non_blocking=True), pe, pl, nz, strength=0.5)
    outs.append(out.images[: len(ids)].clone())
local_out = torch.cat(outs, 0) if outs else torch.empty((0, 1024, 1024, 3), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
t_edit = time.perf_counter() - t0
gathered = sweep.gather_outputs(local_out)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
t_edit_max, t_all_max = sweep.max_over_ranks(t_edit, dev), sweep.max_over_ranks(t_all, dev)
if rank == 0:
    assert gathered.shape[0] == len(entries)
    print(json.dumps({"config": f"PIE-Bench-shaped synthetic sweep, {len(entries)} x 1024x1024, {a.model}, fp16, strength 0.5, micro-batch {B}/GPU",
                      "n_gpus": world, "images": len(entries), "seconds_edit": t_edit_max, "seconds_with_gather": t_all_max,
                      "images_per_s": len(entries) / t_all_max, "gathered_bytes": int(gathered.numel()),
                      "note": f"source images cycle through a pool of {POOL} pre-generated synthetic images per rank (host pinned memory, H2D inside the timed loop); prompts and noises are per image"}), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()

#!/usr/bin/env python
"""Turns one GPU evidence session (scripts/gpu_r2.sh <tag> bench tp full, outputs in gpurun_out/) into the committed files under profiles/:
    python scripts/profile_pack.py <tag>
  <tag>_bench.json, <tag>_tensor_pipe_step.json, <tag>_launches.md, <tag>_{gemm,attn,norm}_full.md, <tag>_gemm_traffic.json"""
import json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
if os.path.exists(f"{G}/bench_{tag}.json"):
    shutil.copy(f"{G}/bench_{tag}.json", f"{P}/{tag}_bench.json")
for x in ("gemm", "attn", "norm"):
    rep = f"{G}/prof_{x}_{tag}.ncu-rep"
    if os.path.exists(rep):
        with open(f"{P}/{tag}_{x}_full.md", "w") as f:
            subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), "full", rep], stdout=f, check=False)
md = f"{P}/{tag}_gemm_full.md"
if os.path.exists(md):
    t = open(md).read()
    reads = [float(x) for x in re.findall(r"dram read: ([0-9.]+) Mbyte", t)]
    writes = [float(x) for x in re.findall(r"dram write: ([0-9.]+) Mbyte", t)]
    durs = [float(x) for x in re.findall(r"duration: ([0-9.]+) us", t)]
    tps = [float(x) for x in re.findall(r"tensor pipe active % \(elapsed\): ([0-9.]+)", t)]
    n = len(reads)
    if n and n == len(writes) == len(durs) == len(tps):
        json.dump({"source": f"ncu --set full --clock-control none, profiles/{tag}_gemm_full.md ({n} consecutive k_gemm_conv launches of one eager SDXL batch-8 edit, scripts/step_probe.py edit)",
                   "dram_bytes_per_launch_avg": sum((r + w) * 1e6 for r, w in zip(reads, writes)) / n, "launches": n,
                   "tensor_pipe_active_pct_time_weighted": sum(d * p for d, p in zip(durs, tps)) / sum(durs),
                   "note": f"a {n}-launch sample; the time-weighted tensor-pipe figure over ALL launches of the step is in {tag}_tensor_pipe_step.json"},
                  open(f"{P}/{tag}_gemm_traffic.json", "w"), indent=1)
tp = f"{G}/{tag}_tensor_pipe_step.json"
if os.path.exists(tp):
    t = json.load(open(tp))
    json.dump(t, open(f"{P}/{tag}_tensor_pipe_step.json", "w"), indent=1)
    e, u = t["edit"], t.get("unet_step")
    tot = e["gpu_time_ms"]
    L = [f"# Launch list of one eager SDXL batch-8 edit ({tag})\n",
         f"`ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed python scripts/step_probe.py edit`: "
         f"{e['launches']} launches, {tot:.1f} ms serialised (cold caches, no PDL overlap; the graph-replayed step in `bench.py` includes gaps and runs under the power cap).\n",
         "| kernel | launches | ms | share | tensor pipe % (time-weighted) |", "|---|---|---|---|---|"]
    for k, v in e["by_kernel"].items():
        L.append(f"| `{k}` | {v['launches']} | {v['ms']:.3f} | {100 * v['ms'] / tot:.1f} % | {v['tensor_pipe_pct']} |")
    L.append(f"\nTensor kernels (`k_gemm_conv`, `k_attention_d64*`): {e['tensor_kernels_launches']} launches, {e['tensor_kernels_time_ms']:.1f} ms = "
             f"{100 * e['tensor_kernels_time_ms'] / tot:.1f} % of the GPU time; time-weighted tensor-pipe activity {t['tensor_pipe_active_pct_time_weighted']:.1f} % over them, "
             f"{t['tensor_pipe_active_pct_all_kernels']:.1f} % over every launch of the edit.")
    if u:
        L.append(f"\nOne ControlNet + UNet evaluation alone (`step_probe.py unet`, CFG batch of 16 rows): {u['launches']} launches, {u['gpu_time_ms']:.1f} ms; tensor-pipe activity "
                 f"{t['unet_step_tensor_pipe_active_pct_tensor_kernels']:.1f} % over its tensor kernels, {t['unet_step_tensor_pipe_active_pct']:.1f} % over all of its launches.")
    open(f"{P}/{tag}_launches.md", "w").write("\n".join(L) + "\n")
print("packed", tag, sorted(f for f in os.listdir(P) if f.startswith(tag)))

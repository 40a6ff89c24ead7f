#!/usr/bin/env python
"""Time-weighted tensor-pipe utilisation over EVERY launch of a profiled range (scripts/step_probe.py), from the ncu CSV:

    python scripts/ncu_tensor_pipe.py gpurun_out/tp_edit.csv [gpurun_out/tp_unet.csv] > profiles/r2_tensor_pipe_step.json

For each kernel: duration and sm__pipe_tensor_cycles_active (% of peak, elapsed).  Weighted by duration this is the fraction of the
step's GPU time during which the tensor pipes were busy — over the tensor-core kernels only, and over all kernels."""
import collections
import csv
import json
import re
import sys

TC = re.compile(r"k_gemm_conv|k_attention_d64|k_attn_vae")


def read(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"kernel": r[ik]})
        v = float(r[iv].replace(",", ""))
        if r[im].startswith("gpu__time_duration"):
            d["ns"] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1.0)
        elif r[im].startswith("sm__pipe_tensor_cycles_active"):
            d["tp"] = v
    return [d for d in per.values() if "ns" in d and "tp" in d]


def summarise(rows):
    tot = sum(d["ns"] for d in rows)
    tc = [d for d in rows if TC.search(d["kernel"])]
    tc_ns = sum(d["ns"] for d in tc)
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in rows:
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", d["kernel"])).replace("fie::", "").replace("void ", "")
        f = fam[name]
        f[0] += 1; f[1] += d["ns"]; f[2] += d["ns"] * d["tp"]
    return {"launches": len(rows), "gpu_time_ms": tot / 1e6, "tensor_kernels_launches": len(tc), "tensor_kernels_time_ms": tc_ns / 1e6,
            "tensor_pipe_active_pct_time_weighted_tensor_kernels": sum(d["ns"] * d["tp"] for d in tc) / tc_ns if tc_ns else None,
            "tensor_pipe_active_pct_time_weighted_all_kernels": sum(d["ns"] * d["tp"] for d in rows) / tot if tot else None,
            "by_kernel": {k: {"launches": v[0], "ms": round(v[1] / 1e6, 3), "tensor_pipe_pct": round(v[2] / v[1], 2) if v[1] else None}
                          for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}}


def main():
    edit = summarise(read(sys.argv[1]))
    out = {"source": f"ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum --clock-control none over ALL "
                     f"{edit['launches']} launches of one eager SDXL batch-8 edit (scripts/step_probe.py edit); per-launch times are serialised "
                     "(no PDL overlap), percentages are of the tensor-pipe peak at the clock of each launch",
           "tensor_pipe_active_pct_time_weighted": edit["tensor_pipe_active_pct_time_weighted_tensor_kernels"],
           "tensor_pipe_active_pct_all_kernels": edit["tensor_pipe_active_pct_time_weighted_all_kernels"], "edit": edit}
    if len(sys.argv) > 2:
        unet = summarise(read(sys.argv[2]))
        out["unet_step"] = unet
        out["unet_step_tensor_pipe_active_pct"] = unet["tensor_pipe_active_pct_time_weighted_all_kernels"]
        out["unet_step_tensor_pipe_active_pct_tensor_kernels"] = unet["tensor_pipe_active_pct_time_weighted_tensor_kernels"]
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()

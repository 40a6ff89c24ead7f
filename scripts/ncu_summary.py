#!/usr/bin/env python
"""Turns ncu outputs (read on the CPU box) into the small text summaries committed under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches_<tag>.csv            > profiles/<tag>_launches.md
    python scripts/ncu_summary.py full gpurun_out/prof_<x>_<tag>.ncu-rep            > profiles/<tag>_<x>_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__cluster_size", "cluster size"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    tot = collections.defaultdict(lambda: [0, 0.0, 1e30, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        us = float(row["Metric Value"]) / 1e3
        t = tot[name]
        t[0] += 1; t[1] += us; t[2] = min(t[2], us); t[3] = max(t[3], us)
        n += 1
    total = sum(v[1] for v in tot.values())
    print(f"# ncu launch list: {n} launches, {total / 1e3:.2f} ms of serialised kernel time (gpu__time_duration.sum, --clock-control none)\n")
    print("| kernel | launches | total ms | share | min us | max us |\n|---|---:|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / total:.1f}% | {v[2]:.1f} | {v[3]:.1f} |")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path.split('/')[-1]} ({len(rows) - 2} captured launches)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## `{d['Kernel Name']}` grid {d['Grid Size']} block {d['Block Size']}\n")
        for key, label in KEYS:
            if key in d and d[key] != "":
                print(f"- {label}: {d[key]} {units[hdr.index(key)]}  (`{key}`)")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

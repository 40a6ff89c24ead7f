#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of each kernel in an ncu report captured with --import-source on.
   python scripts/ncu_hot.py <file.ncu-rep> [top_n] [--nowait]"""
import csv, subprocess, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kern.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
for k in kern:
    h = k["hdr"]; iS = h.index("# Samples"); isrc = h.index("Source"); iex = h.index("Instructions Executed")
    sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iS]) for r in k["rows"])
    agg = collections.Counter()
    for r in k["rows"]:
        for i in sc: agg[h[i][6:]] += int(r[i])
    print("=====", k["name"][:70], "samples", tot, "instrs", len(k["rows"]))
    print("   stall totals:", ", ".join(f"{a}={b}" for a, b in agg.most_common(8)))
    for r in sorted(k["rows"], key=lambda r: -int(r[iS]))[:topn]:
        st = sorted(((h[i][6:], int(r[i])) for i in sc if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"{int(r[iS]):6d} {100*int(r[iS])/max(tot,1):5.1f}%  ex {r[iex]:>8s}  {r[isrc].strip()[:64]:64s} {st}")

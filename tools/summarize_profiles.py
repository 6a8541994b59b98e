#!/usr/bin/env python
"""Turns the raw outputs of tools/profile_round.sh (gpurun_out/<tag>/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r1s2
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
dst = "profiles"
P = lambda n: os.path.join(dst, f"{tag}_{n}")


def last_json_line(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return line
    raise SystemExit(f"no JSON line in {path}")


for n in ("bench.json", "bench_reference.json", "bench_inception.json", "bench_deeplabv3.json", "slide_40k.json"):
    open(P(n), "w").write(last_json_line(os.path.join(src, n)) + "\n")
for n in ("ops.csv", "ops_inception.csv", "ops_deeplabv3.csv", "timeline.txt", "launches.csv"):
    shutil.copy(os.path.join(src, n), P(n))

# ---- launch list -> per-kernel shares
rows = list(csv.reader(open(os.path.join(src, "launches.csv"))))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(P("launch_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,share\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{us:.1f},{us / tot:.3f}\n")
    f.write(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.1f},1.000\n")

# ---- ncu --set full raw pages -> one markdown table
WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1/TEX(+smem) % of peak", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("instructions", "smsp__inst_executed.sum"),
    ("smem LSU wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
]
cols = []
for rep in ("full_dec", "full_dl16", "full_dl8"):
    path = os.path.join(src, rep + ".raw.csv")
    if not os.path.exists(path):
        continue
    rr = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rr) if r and r[0] == "ID"][0]
    names, units = rr[h], rr[h + 1]
    for r in rr[h + 2:]:
        if len(r) < len(names):
            continue
        d = {}
        for n, u, v in zip(names, units, r):
            d[n] = (v, u)
        cols.append(d)
with open(P("ncu_top_kernels.md"), "w") as f:
    f.write(f"# ncu --set full, top kernels of one forward step (batch 32), {tag}\n\n"
            "Captured by tools/profile_round.sh with `ncu --set full --clock-control none --import-source on` (direct\n"
            "launches, `--no-graph`), one launch each; cold-cache, serialised -- durations are for shares, not bench\n"
            "values.  Columns: launches 75/76/77 (dec9b, dec10a, dec10b) of the profiled step, one fused dense layer of\n"
            "block 2 (64x64 maps) and one of block 4 (16x16 maps).\n\n")
    heads = []
    for d in cols:
        k = re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("void ", "")
        heads.append(k)
    f.write("| metric | " + " | ".join(heads) + " |\n|---|" + "---|" * len(heads) + "\n")
    for label, key in WANT:
        vals = []
        for d in cols:
            hit = [k for k in d if k == key] or [k for k in d if k.endswith("." + key) and d[k][0]] or \
                  [k for k in d if k.endswith(key) and d[k][0]]
            vals.append(f"{d[hit[0]][0]} {d[hit[0]][1]}".strip() if hit else "n/a")
        f.write(f"| {label} (`{key}`) | " + " | ".join(vals) + " |\n")
print("wrote", sorted(n for n in os.listdir(dst) if n.startswith(tag)))

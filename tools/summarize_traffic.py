"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch of ONE forward step)
-> profiles/r2_step_traffic.json + a per-kernel-family table.  Usage: summarize_traffic.py in.csv out.json out.md"""
import csv, json, sys, collections
src, out_json, out_md = sys.argv[1:4]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(lambda: collections.defaultdict(float))   # (id) -> metric -> value
names = {}
unit = {}
for r in rows[1:]:
    if len(r) < len(hdr) or not r[ix["ID"]].isdigit():
        continue
    kid = int(r[ix["ID"]])
    names[kid] = r[ix["Kernel Name"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
            "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1)
    per[kid][r[ix["Metric Name"]]] += v * mult
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot_r = tot_w = tot_t = 0.0
for kid, m in per.items():
    nm = names[kid].split("(")[0]
    rd, wr, t = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 0)
    f = fam[nm]; f[0] += 1; f[1] += rd; f[2] += wr; f[3] += t
    tot_r += rd; tot_w += wr; tot_t += t
alg = 32 * (196608 + 262144)
note = (f"dram__bytes_read.sum + dram__bytes_write.sum over the {len(per)} launches of one batch-32 forward (ncu, direct "
        f"launches, serialised, cold-cache per kernel): {tot_r/1e6:.1f} MB read + {tot_w/1e6:.1f} MB written; algorithmic "
        f"bytes per step {alg/1e6:.1f} MB (tiles in + probabilities out), weights 34.6 MB once")
json.dump({"dram_bytes_per_step": tot_r + tot_w, "dram_read": tot_r, "dram_write": tot_w, "launches": len(per),
           "serialised_kernel_time_s": tot_t, "algorithmic_bytes_per_step": alg, "note": note}, open(out_json, "w"), indent=1)
with open(out_md, "w") as f:
    f.write("# Whole-step DRAM traffic of the batch-32 DenseNet U-Net forward (ncu)\n\n" + note + "\n\n")
    f.write("| kernel | launches | DRAM read MB | DRAM written MB | time us (serialised) | GB/s |\n|---|---|---|---|---|---|\n")
    for nm, (n, rd, wr, t) in sorted(fam.items(), key=lambda kv: -kv[1][3]):
        f.write(f"| {nm[:70]} | {n} | {rd/1e6:.1f} | {wr/1e6:.1f} | {t*1e6:.1f} | {(rd+wr)/t/1e9 if t else 0:.0f} |\n")
print(note)

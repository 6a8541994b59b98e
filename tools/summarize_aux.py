"""ncu CSV of tools/aux_kernels.py -> markdown table (second occurrence of each kernel = warm run)."""
import csv, sys, collections, json
src, out_md = sys.argv[1:3]
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if len(sys.argv) < 4 else float(sys.argv[3])
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.OrderedDict()
for r in rows[1:]:
    if not r[ix["ID"]].isdigit():
        continue
    k = int(r[ix["ID"]])
    d = per.setdefault(k, {"name": r[ix["Kernel Name"]].split("(")[0]})
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
            "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}.get(u, 1)
    d[r[ix["Metric Name"]]] = v * mult
seen = collections.Counter()
lines = ["| kernel (warm launch) | time us | DRAM read MB | DRAM written MB | DRAM GB/s | frac of measured HBM peak | L2 traffic MB |",
         "|---|---|---|---|---|---|---|"]
n_total = collections.Counter(d["name"] for d in per.values())
for d in per.values():
    seen[d["name"]] += 1
    if seen[d["name"]] <= n_total[d["name"]] // 2:      # first half = cold pass
        continue
    t = d.get("gpu__time_duration.sum", 0); rd = d.get("dram__bytes_read.sum", 0); wr = d.get("dram__bytes_write.sum", 0)
    gbs = (rd + wr) / t / 1e9 if t else 0
    lines.append(f"| {d['name'][:60]} | {t*1e6:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {gbs:.0f} | {gbs/peak:.2f} | {d.get('lts__t_bytes.sum',0)/1e6:.1f} |")
open(out_md, "w").write("# HBM-bound helper kernels under ncu (tools/aux_kernels.py; peak = MEASURED_PEAKS hbm_gbs %.0f GB/s)\n\n" % peak + "\n".join(lines) + "\n")
print("\n".join(lines))

#!/usr/bin/env bash
# Round-end evidence on one B200: full GPU suite, smoke, bench (both arms, precision modes, other ensemble members),
# config 5, lanes timeline, ncu launch list + one ncu --set full capture.   Usage (on the box): tools/call_final.sh <tag>
set -u
TAG=${1:-r2final}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee -a "$OUT/summary.txt"
grep -E "passed|failed|FAILED" "$OUT/pytest_gpu.log" | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/smoke.log"
timeout 600 python bench.py --dump-ops "$OUT/ops.csv" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"
for P in fp32 tf32x3; do
  timeout 300 python bench.py --precision $P --steps 10 --no-slide --no-parity --no-cpu-baseline > "$OUT/bench_$P.json" 2>> "$OUT/bench.err"; echo "bench $P rc=$?" | tee -a "$OUT/summary.txt"
done
for M in inception deeplabv3; do
  timeout 300 python bench.py --model $M --steps 60 --no-slide --no-parity --no-cpu-baseline > "$OUT/bench_$M.json" 2>> "$OUT/bench.err"; echo "bench $M rc=$?" | tee -a "$OUT/summary.txt"
done
timeout 400 python bench.py --workload config5 --slide 8192 --steps 2 > "$OUT/config5_n1.json" 2>> "$OUT/bench.err"; echo "config5 rc=$?" | tee -a "$OUT/summary.txt"
timeout 100 python tools/lanes_timeline.py 3 > "$OUT/lanes_timeline.txt" 2>&1
timeout 100 python tests/stamp_ops.py > "$OUT/timeline.txt" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --lanes 1 --no-slide --no-parity --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
true
true
true
python - <<PY
import json
for f in ("bench", "bench_fp32", "bench_tf32x3", "bench_inception", "bench_deeplabv3", "config5_n1", "bench_reference"):
    try:
        d = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "e2e", d.get("e2e", {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"),
              "single", (d.get("single_stream") or {}).get("value"), "slide", (d.get("slide") or {}).get("tiles_per_s"))
    except Exception as e:
        print(f, "no line", e)
PY
tail -3 "$OUT/bench.err"

#!/usr/bin/env bash
# N-GPU evidence: default bench line (forward weak scaling + sharded 40k slide) and BASELINE configs[4] (ensemble + CRF).
set -u
TAG=${1:-r2scale}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench n$N rc=$?" | tee -a "$OUT/summary.txt"
timeout 600 $TR bench.py --gpus $N --workload config5 --slide 16384 --steps 2 > "$OUT/config5_n$N.json" 2> "$OUT/config5_n$N.err"; echo "config5 n$N rc=$?" | tee -a "$OUT/summary.txt"
python - <<PY
import json
for f in ("$OUT/bench_n$N.json", "$OUT/config5_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), json.dumps(d.get("slide") or d.get("config5"))[:900])
    except Exception as e:
        print(f, "no line", e)
PY
for f in "$OUT"/*.err; do tail -n 2 "$f"; done

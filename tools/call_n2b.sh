#!/usr/bin/env bash
set -u
TAG=${1:-r2n2b}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee -a "$OUT/summary.txt"
grep -E "passed|failed" "$OUT/pytest_gpu.log" | tail -2
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench n$N rc=$?" | tee -a "$OUT/summary.txt"
cut -c1-300 "$OUT/bench_n$N.json"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
    print(json.dumps(d.get("slide"),indent=1))
except Exception as e: print("no line", e)
PY
tail -5 "$OUT/bench_n$N.err"

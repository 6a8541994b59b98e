#!/usr/bin/env bash
# One GPU call: lane tests + bench lines with 1 / 3 / 4 lanes.  Usage (on the box): tools/call_lanes.sh <tag>
set -u
tag=${1:-r2x2}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_properties.py -q -x -k "lanes or sharded or deterministic" > $out/pytest_lanes.log 2>&1
tail -5 $out/pytest_lanes.log
for L in 3 4; do
  timeout 600 python bench.py --lanes $L --steps 200 --warmup 5 --no-parity > $out/bench_l$L.json 2> $out/bench_l$L.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_l$L.json").read().strip().splitlines()[-1])
    print("lanes $L value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_stream"]["value"]),
          "frac", round(d["roofline"]["frac"], 3), "slide", d.get("slide", {}).get("tiles_per_s"), d.get("slide", {}).get("seconds"))
except Exception as e:
    print("lanes $L: no line", e); print(open("$out/bench_l$L.err").read()[-2000:])
PY
done

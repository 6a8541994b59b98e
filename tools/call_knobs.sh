#!/usr/bin/env bash
# Planner knobs under lanes: short bench runs (no slide / parity / CPU legs).  Usage (on the box): tools/call_knobs.sh <tag>
set -u
tag=${1:-r2x3}
out=gpurun_out/$tag
mkdir -p $out
run() {  # name, lanes, env...
  name=$1; lanes=$2; shift 2
  env "$@" timeout 300 python bench.py --lanes $lanes --steps 150 --warmup 5 --no-parity --no-slide --no-cpu-baseline \
      > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"]), "single", round(d["single_stream"]["value"]), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name: no line", e); print(open("$out/$name.err").read()[-1500:])
PY
}
run base3 3 A=1
run base2 2 A=1
run base5 5 A=1
run db0 3 DP_DENSE_BLOCK=0
run db2 3 DP_DENSE_BLOCK=2
run rh16 3 DP_DL_MIN_ITEMS16=64
run rh16_db0 3 DP_DL_MIN_ITEMS16=64 DP_DENSE_BLOCK=0
run rh16b 3 DP_DL_MIN_ITEMS16=16
run split2 2 DP_SPLIT=2
run subitems 3 DP_D_SUB_BY_ITEMS=1
run sub2 3 DP_PREFER_SUB2=1
run nophase 3 DP_NO_PHASE_SPLIT=1

#!/usr/bin/env bash
set -u
TAG=${1:-r2s3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee -a "$OUT/summary.txt"
grep -E "passed|failed|error" "$OUT/pytest_gpu.log" | tail -3
timeout 600 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
   --launch-skip 156 --launch-count 78 --csv --log-file "$OUT/step_traffic.csv" python tools/one_step.py > "$OUT/one_step.log" 2>&1; echo "ncu traffic rc=$?" | tee -a "$OUT/summary.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/smoke.log" | tail -3; cut -c1-1500 "$OUT/bench.json"; tail -5 "$OUT/bench.err"

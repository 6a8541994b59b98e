#!/usr/bin/env bash
# Round-end evidence on one B200: full GPU suite, smoke, bench (both arms), launch list + one ncu --set full capture.
set -u
TAG=${1:-r2final}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee -a "$OUT/summary.txt"
grep -E "passed|failed|FAILED" "$OUT/pytest_gpu.log" | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
timeout 600 python bench.py --dump-ops "$OUT/ops.csv" > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --precision fp32 --steps 10 --no-slide --no-parity --no-cpu-baseline > "$OUT/bench_fp32.json" 2>> "$OUT/bench.err"; echo "bench fp32 rc=$?" | tee -a "$OUT/summary.txt"
timeout 100 python tests/stamp_ops.py > "$OUT/timeline.txt" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-slide --no-parity --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'dense_block_kernel|conv_tc_kernel' --launch-skip 32 --launch-count 6 -o "$OUT/top_kernels" python tools/one_step.py > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?" | tee -a "$OUT/summary.txt"
ncu -i "$OUT/top_kernels.ncu-rep" --page raw --csv > "$OUT/top_kernels_raw.csv" 2>/dev/null
cut -c1-400 "$OUT/bench.json"; tail -3 "$OUT/bench.err"

#!/usr/bin/env bash
# First GPU call of the next round (one B200): everything that was written after round 1's GPU budget ran out.
# Every step that could hang sits under its own `timeout`.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1500 -- 'bash tools/first_call_next_round.sh r2s1'
set -u
TAG=${1:-r2s1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
# 1. the regular GPU suite (includes the late-written CRF helper / local_part / world-of-one / ingest tests)
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1; echo "pytest -m gpu rc=$?" | tee -a "$OUT/summary.txt"
# 2. cta_group::2 MMA probe
( cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe_cta2 mma_probe_cta2.cu ) > "$OUT/probe_build.log" 2>&1
timeout 60 tools/mma_probe_cta2 > "$OUT/mma_probe_cta2.txt" 2>&1; echo "cta2 probe rc=$?" | tee -a "$OUT/summary.txt"
timeout 600 env DP_TEST_UNVERIFIED=1 python -m pytest tests/test_gpu_wsi_ingest.py -m gpu -q > "$OUT/pytest_ingest.log" 2>&1
echo "nvJPEG ingest tests rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python bench.py --steps 50 --warmup 5 --dump-ops "$OUT/ops.csv" > "$OUT/bench.json" 2> "$OUT/bench.err"
for mode in 1 2; do
  echo "bench DP_DL_STACK=$mode rc=$?" | tee -a "$OUT/summary.txt"
done
# 3b. role timelines of one 16x16 and one 8x8 dense layer (what bounds conv4 / conv5: L2 -> SM streaming or the
timeout 120 python tests/trace_ops.py 36 58 > "$OUT/trace_conv4_conv5.txt" 2>&1
# 4. slide-level run (tissue mask, grid, forward, stitch) -- the sharded variant needs `gpurun --gpus 2`:
#    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload slide --slide 16384 --steps 2
timeout 300 python bench.py --workload slide --slide 16384 --steps 2 > "$OUT/slide_16k_n1.json" 2> "$OUT/slide.err"

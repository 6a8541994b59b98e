#!/usr/bin/env bash
# 2-GPU validation of the x-stripe sharded slide run (gpurun --gpus 2 -- 'bash tools/call_n2.sh TAG')
set -u
TAG=${1:-r2n2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > "$OUT/pytest_multi.log" 2>&1; echo "test_gpu_multi rc=$?" | tee -a "$OUT/summary.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --workload slide --slide 16384 --steps 2 > "$OUT/slide_16k_n2.json" 2> "$OUT/slide_16k_n2.err"; echo "slide16k n2 rc=$?" | tee -a "$OUT/summary.txt"
timeout 400 $TR bench.py --gpus 2 --workload slide --slide 40000 --tta FLIP_LEFT_RIGHT,ROTATE_90,ROTATE_180 --steps 1 > "$OUT/slide_40k_n2.json" 2> "$OUT/slide_40k_n2.err"; echo "slide40k n2 rc=$?" | tee -a "$OUT/summary.txt"
tail -n 5 "$OUT/pytest_multi.log"; cut -c1-600 "$OUT/slide_16k_n2.json" "$OUT/slide_40k_n2.json"; tail -n 5 "$OUT"/*.err

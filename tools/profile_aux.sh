#!/usr/bin/env bash
# ncu capture of the HBM-bound helper kernels (one GPU): gpurun -- 'bash tools/profile_aux.sh TAG'
set -u
TAG=${1:-r2aux}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
  --clock-control none -k regex:'stitch|finalize|stem_s2d|pyramid|tissue|morph|maxpool|bn_act_pool' --csv \
  --log-file "$OUT/aux_kernels.csv" python tools/aux_kernels.py > "$OUT/aux.log" 2>&1
echo "ncu rc=$?"; tail -2 "$OUT/aux.log"; wc -l "$OUT/aux_kernels.csv"

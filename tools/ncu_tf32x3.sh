#!/usr/bin/env bash
# ncu --set full of the 3xTF32 halo kernel on the decoder layers of a batch-32 forward (third of three direct-launch
# forwards: 51 halo launches per forward (42 dense-layer 3x3 convs on maps >= 16x16 + 9 decoder convs), the last 9 are dec6b..dec10b) and of the generic kernel on the transition convs.
set -u
OUT=gpurun_out/${1:-r2x18}; mkdir -p $OUT
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_halo_tf32x3_kernel --launch-skip 144 --launch-count 9 -o $OUT/halo python tools/one_step.py tf32x3 > $OUT/ncu_halo.log 2>&1; echo "ncu halo rc=$?"
ncu -i $OUT/halo.ncu-rep --page raw --csv > $OUT/halo_raw.csv 2>/dev/null
rm -f $OUT/halo.ncu-rep
tail -2 $OUT/ncu_halo.log

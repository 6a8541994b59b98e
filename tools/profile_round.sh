#!/usr/bin/env bash
# Runs the measurement set whose summaries are committed under profiles/ (one B200).  Usage: tools/profile_round.sh TAG
set -u
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
KRE='regex:conv_tc|dense_layer|stem_s2d|bn_act|maxpool'
python bench.py --steps 50 --warmup 5 --dump-ops "$OUT/ops.csv" > "$OUT/bench.json" 2> "$OUT/bench.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2>> "$OUT/bench.err"
python bench.py --model inception --steps 30 --warmup 5 --dump-ops "$OUT/ops_inception.csv" > "$OUT/bench_inception.json" 2>> "$OUT/bench.err"
python bench.py --model deeplabv3 --steps 30 --warmup 5 --dump-ops "$OUT/ops_deeplabv3.csv" > "$OUT/bench_deeplabv3.json" 2>> "$OUT/bench.err"
python tests/stamp_ops.py > "$OUT/timeline.txt" 2>&1
# launch list of ONE forward step (the 4th: 3 warm-up forwards x 78 launches are skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 234 -c 78 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/ncu_launches.log" 2>&1
# full captures: dec9b / dec10a / dec10b (launches 75..77 of that step), one dense layer of block 2 and one of block 4
ncu --set full --clock-control none --import-source on -k "$KRE" -s 309 -c 3 -o "$OUT/full_dec" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/ncu_full_dec.log" 2>&1
ncu --set full --clock-control none --import-source on -k "$KRE" -s 239 -c 1 -o "$OUT/full_dl16" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/ncu_full_dl16.log" 2>&1
ncu --set full --clock-control none --import-source on -k "$KRE" -s 270 -c 1 -o "$OUT/full_dl8" -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/ncu_full_dl8.log" 2>&1
for r in full_dec full_dl16 full_dl8; do
  ncu -i "$OUT/$r.ncu-rep" --page raw --csv > "$OUT/$r.raw.csv" 2>/dev/null
done
python bench.py --workload slide --slide 40000 --steps 1 --tta FLIP_LEFT_RIGHT,ROTATE_90,ROTATE_180 > "$OUT/slide_40k.json" 2> "$OUT/slide.err"
tail -c 600 "$OUT/bench.json"; echo; cut -c1-160 "$OUT/bench_inception.json" "$OUT/bench_deeplabv3.json" "$OUT/slide_40k.json"
ls -la "$OUT"

#!/bin/bash
# Developer tool: everything profiles/ is built from, in one GPU call.  Usage: tools/capture_profiles.sh TAG
# Writes gpurun_out/{bench,bench_ref,bench_english}_TAG.json, launches_TAG.csv, prof_TAG.ncu-rep, raw_TAG.csv.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
python bench.py --workload english --no-cpu --steps 10 --warmup 3 > $OUT/bench_english_$TAG.json 2>> $OUT/bench_$TAG.err
# the same command without and then with ncu (launch list: every kernel of warm-up + 2 timed steps + the e2e and decode legs)
python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
# full capture of the main kernels of one device-resident step (the 4th: after three warm-up steps)
K='regex:^(pretok_kernel|lookup_kernel|lanemerge_kernel|emit_kernel)$'
ncu --set full --clock-control none --import-source on --kernel-name "$K" --launch-skip 36 --launch-count 12 \
    -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
tail -2 $OUT/ncu_full_$TAG.log

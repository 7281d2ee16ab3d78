#!/bin/bash
# Developer tool: full GPU test suite, then the bench lines and the ncu captures of tools/capture_profiles.sh (without
# the reference arm).  Usage: tools/final_capture.sh TAG
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest_$TAG.log
grep -q " passed" $OUT/pytest_$TAG.log && ! grep -q "failed" $OUT/pytest_$TAG.log || exit 1
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || exit 1
python bench.py --workload english --no-cpu --steps 10 --warmup 3 > $OUT/bench_english_$TAG.json 2>> $OUT/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_launches_$TAG.log 2>&1
K='regex:^(pretok_kernel|lookup_kernel|lanemerge_kernel|emit_kernel)$'
ncu --set full --clock-control none --import-source on --kernel-name "$K" --launch-skip 36 --launch-count 12 \
    -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
tail -1 $OUT/ncu_full_$TAG.log

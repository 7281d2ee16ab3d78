#!/bin/bash
# round 2, final capture on one B200: full GPU test suite, smoke, one bench line per BASELINE config (+ stored-pattern
# split, reference arm, latency table), ncu launch list and --set full captures (encode step, decode) of the same
# command, each only after the plain run of that command exited 0.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.txt
tail -4 $O/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_smoke.txt 2>&1; tail -1 $O/r02_smoke.txt
b() { n=$1; shift; timeout 1200 python bench.py "$@" > $O/r02_bench_$n.json 2> $O/r02_bench_$n.err; echo "$n rc=$? $(head -c 160 $O/r02_bench_$n.json)"; }
b mixed --steps 10 --warmup 3 --latency
b reference --impl reference --steps 3 --warmup 1
b english1m --workload english1m --steps 10 --warmup 3
b single1g --workload single1g --steps 5 --warmup 3 --no-cpu
b adversarial --workload adversarial --steps 5 --warmup 3
b roundtrip8 --workload roundtrip64g --shards 8 --no-cpu
b mixed_cfgsplit --split config --steps 5 --warmup 3 --no-cpu --quick
b english --workload english --steps 5 --warmup 3 --no-cpu --quick
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-e2e"
timeout 600 $CMD > $O/r02_plain.json 2> $O/r02_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $CMD > $O/r02_ncu_launches.log 2>&1
K='regex:^(pretok_kernel|lookup_kernel|lanemerge_kernel|emit_kernel|longmerge_block_kernel|longmerge_warp_kernel)'
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name "$K" --launch-skip 45 --launch-count 15 -o $O/r02_prof_encode -f $CMD > $O/r02_ncu_full_encode.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name 'regex:decode_gather|decode_validate' --launch-skip 2 --launch-count 2 -o $O/r02_prof_decode -f $CMD > $O/r02_ncu_full_decode.log 2>&1
ls -la $O/r02_prof_*.ncu-rep

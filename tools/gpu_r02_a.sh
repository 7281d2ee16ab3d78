#!/bin/bash
# round 2, call A: GPU tests + one bench line per BASELINE config with the round-1 kernels (+ L2 window A/B)
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/r02a_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --latency > $O/r02a_bench_mixed.json 2> $O/r02a_bench_mixed.err
TEKKEN_B200_L2PIN=0 timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-cpu > $O/r02a_bench_mixed_nopin.json 2> $O/r02a_bench_mixed_nopin.err
timeout 300 python bench.py --workload english1m --steps 20 > $O/r02a_bench_english1m.json 2> $O/r02a_bench_english1m.err
timeout 600 python bench.py --workload adversarial --steps 3 > $O/r02a_bench_adversarial.json 2> $O/r02a_bench_adversarial.err
timeout 900 python bench.py --workload single1g --steps 5 > $O/r02a_bench_single1g.json 2> $O/r02a_bench_single1g.err
timeout 900 python bench.py --workload roundtrip64g --shards 8 > $O/r02a_bench_roundtrip8.json 2> $O/r02a_bench_roundtrip8.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02a_bench_reference.json 2> $O/r02a_bench_reference.err
tail -3 $O/r02a_pytest.log
for f in mixed mixed_nopin english1m adversarial single1g roundtrip8 reference; do echo "== $f"; head -c 600 $O/r02a_bench_$f.json; echo; tail -2 $O/r02a_bench_$f.err; done

#!/bin/bash
# round 2, call B: config-pattern split on the GPU (tests + bench line)
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02b_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_config_pattern.py -m gpu -x -q > $O/r02b_pytest_cfg.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest_cfg.log
timeout 600 python bench.py --split config --steps 5 --warmup 3 --quick > $O/r02b_bench_mixed_cfg.json 2> $O/r02b_bench_mixed_cfg.err
tail -30 $O/r02b_pytest_cfg.log
head -c 1500 $O/r02b_bench_mixed_cfg.json; tail -3 $O/r02b_bench_mixed_cfg.err

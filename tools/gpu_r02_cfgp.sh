#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02cfgp_build.log 2>&1
CMD="python bench.py --split config --steps 1 --warmup 3 --no-cpu --quick --no-e2e"
timeout 600 $CMD > $O/r02cfgp_plain.json 2> $O/r02cfgp_plain.err || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:cfg_walk --launch-skip 3 -c 1 -o $O/r02cfgp_walk -f $CMD > $O/r02cfgp_ncu.log 2>&1
ls -la $O/r02cfgp_walk.ncu-rep

#!/bin/bash
# round 2, call W: full ncu capture of the decode kernels
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02w_build.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > $O/r02w_plain.json 2> $O/r02w_plain.err || exit 1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"decode_gather|decode_validate" --launch-skip 2 -c 2 -o $O/r02w_decode python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > $O/r02w_ncu.log 2>&1
ls -la $O/r02w_decode.ncu-rep

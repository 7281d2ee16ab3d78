#!/bin/bash
# round 2, call H: experiments -- concurrent sub-batch pipelines on one GPU; D2H rate by pinned-allocation flag
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02h_build.log 2>&1
timeout 600 python tools/overlap_probe.py > $O/r02h_overlap.json 2> $O/r02h_overlap.err; cat $O/r02h_overlap.json; tail -n 3 $O/r02h_overlap.err
timeout 300 python tools/d2h_probe.py > $O/r02h_d2h.json 2> $O/r02h_d2h.err; cat $O/r02h_d2h.json; tail -n 3 $O/r02h_d2h.err

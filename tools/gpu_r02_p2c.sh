#!/bin/bash
# round 2: does the packed-id download help or hurt with two processes on the box?  (same box, back to back)
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r02p2c_nproc.txt
python -c "import __graft_entry__ as g; g.build()" > $O/r02p2c_build.log 2>&1
for m in 0 -1 0 -1; do
TEKKEN_B200_PACK_IDS=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --quick > $O/r02p2c_n2_$m.json 2> $O/r02p2c_n2_$m.err
python - $m <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02p2c_n2_%s.json"%sys.argv[1])); print("pack", sys.argv[1], "value", round(j["value"],1), "e2e", round(j["e2e"]["value"],2), round(j["e2e"]["ms_per_step"],2))
except Exception as e: print("ERR", e)
PY
done
cat $O/r02p2c_nproc.txt

#!/bin/bash
# round 2: last check of the committed tree -- the default bench invocation, timed
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02last_build.log 2>&1
S=$(date +%s)
python bench.py > $O/r02last_bench.json 2> $O/r02last_bench.err; echo "bench rc=$?"
echo "wall seconds: $(( $(date +%s) - S ))"
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02last_bench.json")); print(round(j["value"],2), round(j["ms_per_step"],3), round(j["e2e"]["value"],2), j["gpu_launches"], j["roofline"]["frac"], j["cpu_baseline"]["value"])
PY
S=$(date +%s)
python bench.py --impl reference > $O/r02last_ref.json 2> $O/r02last_ref.err; echo "ref rc=$?  wall seconds: $(( $(date +%s) - S ))"
head -c 200 $O/r02last_ref.json

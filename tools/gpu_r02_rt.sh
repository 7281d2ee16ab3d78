#!/bin/bash
# round 2: BASELINE configs[4] at full size (64 shards) on ONE GPU -- the id checksum must equal the 8-GPU run's
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02rt_build.log 2>&1
timeout 600 python bench.py --workload roundtrip64g --shards 8 --no-cpu > $O/r02rt_roundtrip8.json 2> $O/r02rt_roundtrip8.err
timeout 1500 python bench.py --workload roundtrip64g --shards 64 --no-cpu > $O/r02rt_roundtrip64.json 2> $O/r02rt_roundtrip64.err
python - <<'PY'
import json
for n in ("roundtrip8","roundtrip64"):
    try:
        j=json.load(open("gpurun_out/r02rt_%s.json"%n)); print(n, round(j["value"],2), j["encode"], j["decode"], j["e2e"]["seconds"], j["checks"])
    except Exception as e: print(n, "ERR", e)
PY

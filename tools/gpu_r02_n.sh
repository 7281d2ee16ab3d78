#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02n_build.log 2>&1
timeout 600 python bench.py --workload adversarial --steps 2 --no-cpu > $O/r02n_bench_adversarial.json 2> $O/r02n_bench_adversarial.err
timeout 600 python bench.py --workload adversarial --pieces 7 --steps 2 --no-cpu > $O/r02n_bench_adversarial7.json 2> $O/r02n_bench_adversarial7.err
python - <<'PY'
import json
for f in ("bench_adversarial","bench_adversarial7"):
    try:
        j=json.load(open("gpurun_out/r02n_%s.json"%f))
        print(f, round(j["value"],3), "ms", round(j["ms_per_step"],4), j.get("long_piece_stage"), j["roofline"]["stage_ms"]["longmerge"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/r02n_bench_adversarial.err

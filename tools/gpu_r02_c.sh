#!/bin/bash
# round 2, call C: single-block latency path (tests + latency table)
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02c_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
timeout 300 python bench.py --workload english1m --steps 10 --latency --no-cpu > $O/r02c_bench_english1m.json 2> $O/r02c_bench_english1m.err
tail -30 $O/r02c_pytest.log
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02c_bench_english1m.json"))
for r in j.get("latency",[]): print(r)
PY
tail -3 $O/r02c_bench_english1m.err

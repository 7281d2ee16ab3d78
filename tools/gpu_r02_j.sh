#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02j_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02j_pytest.log
tail -8 $O/r02j_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02j_bench_mixed.json 2> $O/r02j_bench_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02j_bench_mixed.json")); print(round(j["value"],2), j["e2e"])
PY
tail -n 3 $O/r02j_bench_mixed.err

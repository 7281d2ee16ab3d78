#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02i_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine or file or fuzz_batch or config2_mixed" > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest.log
tail -5 $O/r02i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02i_bench_mixed.json 2> $O/r02i_bench_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02i_bench_mixed.json")); print(round(j["value"],2), j["e2e"])
PY
tail -n 3 $O/r02i_bench_mixed.err

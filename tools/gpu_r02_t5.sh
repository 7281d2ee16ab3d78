#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t5_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -k "decode or latency or golden or concurrent or basic" > $O/r02t5_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t5_pytest.log
tail -12 $O/r02t5_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 3 --no-cpu --quick --no-e2e --latency > $O/r02t5_mixed.json 2> $O/r02t5_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t5_mixed.json"))
for r in j["latency"]: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
PY
tail -3 $O/r02t5_mixed.err

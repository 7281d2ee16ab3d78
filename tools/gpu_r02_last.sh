#!/bin/bash
# round 2: last check of the committed tree -- full GPU suite, smoke, the default bench invocation
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02last_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/r02last_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02last_pytest.log
tail -3 $O/r02last_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
TEKKEN_B200_TRACE=1 python bench.py --no-cpu --quick > $O/r02last_bench.json 2> $O/r02last_bench.err; echo "bench rc=$?"
grep "came back as" $O/r02last_bench.err | tail -2
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02last_bench.json")); print(round(j["value"],2), round(j["ms_per_step"],3), round(j["e2e"]["value"],2), j["gpu_launches"], j["decode"]["roundtrip_byte_exact"])
PY

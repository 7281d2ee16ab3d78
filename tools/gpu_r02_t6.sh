#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t6_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_debug_bounds.py tests/test_gpu_parity.py -m gpu -q -x -k "debug or regular_build or decode or latency" > $O/r02t6_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t6_pytest.log
tail -15 $O/r02t6_pytest.log | cut -c1-400
timeout 600 python bench.py --steps 3 --no-cpu --quick --no-e2e > $O/r02t6_mixed.json 2> $O/r02t6_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t6_mixed.json")); d=j["decode"]
print("encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), d["roundtrip_byte_exact"])
PY

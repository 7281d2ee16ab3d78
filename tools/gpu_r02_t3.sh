#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t3_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -k "decode or roundtrip or config5 or engine or golden" > $O/r02t3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t3_pytest.log
tail -4 $O/r02t3_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 5 --no-cpu --quick --no-e2e > $O/r02t3_mixed.json 2> $O/r02t3_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t3_mixed.json")); d=j["decode"]
print("encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), "GB/s", round(d["value"],1), "frac", round(d["hbm_frac"],4), d["roundtrip_byte_exact"])
PY

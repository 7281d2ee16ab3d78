#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02cfg3_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_config_pattern.py -m gpu -q -x > $O/r02cfg3_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02cfg3_pytest.log
tail -3 $O/r02cfg3_pytest.log | cut -c1-200
timeout 600 python bench.py --split config --steps 5 --warmup 3 --no-cpu --quick > $O/r02cfg3_mixed_cfgsplit.json 2> $O/r02cfg3_mixed_cfgsplit.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02cfg3_mixed_cfgsplit.json")); st=j["roofline"]["stage_ms"]
print("cfg split ms", round(j["ms_per_step"],3), "GB/s", round(j["value"],2), {k:v for k,v in st.items() if k.startswith("pretok")}, j["e2e"]["ms_per_step"])
PY

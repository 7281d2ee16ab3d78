#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t4_build.log 2>&1
timeout 600 python bench.py --steps 5 --no-cpu --quick --no-e2e > $O/r02t4_mixed.json 2> $O/r02t4_mixed.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t4_mixed.json")); st=j["roofline"]["stage_ms"]
print("encode ms", round(j["ms_per_step"],3), "lookup", st["lookup"], "pretok", st["pretok"], "emit", st["emit"], "lanemerge", round(sum(v for k,v in st.items() if k.startswith("lanemerge")),3))
PY
timeout 600 python bench.py --workload english --steps 5 --no-cpu --quick --no-e2e > $O/r02t4_english.json 2> $O/r02t4_english.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02t4_english.json")); st=j["roofline"]["stage_ms"]
print("english ms", round(j["ms_per_step"],3), "lookup", st["lookup"])
PY
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02t4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t4_pytest.log
tail -4 $O/r02t4_pytest.log | cut -c1-300

#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t8_build.log 2>&1
for w in mixed english single1g; do
timeout 600 python bench.py --workload $w --steps 5 --no-cpu --quick --no-e2e > $O/r02t8_$w.json 2> $O/r02t8_$w.err
python - $w <<'PY'
import json,sys
j=json.load(open("gpurun_out/r02t8_%s.json"%sys.argv[1])); st=j["roofline"]["stage_ms"]
print(sys.argv[1], "ms", round(j["ms_per_step"],3), "pretok", st["pretok"], "carry", st["pretok_carry"], "lookup", st["lookup"])
PY
done
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02t8_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t8_pytest.log
tail -4 $O/r02t8_pytest.log | cut -c1-300

#!/bin/bash
# round 2, call Z2: packed-id download after tuning -- engine tests, e2e A/B, config 3 through tk_encode
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02z2_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "engine or all_visible or config3 or file" > $O/r02z2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02z2_pytest.log
tail -4 $O/r02z2_pytest.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 900 python bench.py --steps 8 --no-cpu > $O/r02z2_$n.json 2> $O/r02z2_$n.err
  python - "$n" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02z2_%s.json"%sys.argv[1])); e=j["e2e"]
    print(sys.argv[1], "dev ms", round(j["ms_per_step"],3), "e2e ms", round(e["ms_per_step"],2), "GB/s", round(e["value"],2), "pageable", e.get("pageable_input",{}).get("ms_per_step"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
run auto TEKKEN_B200_PACK_IDS=-1
run off TEKKEN_B200_PACK_IDS=0
run auto_t8 TEKKEN_B200_PACK_IDS=-1 TEKKEN_B200_COPY_THREADS=8
TEKKEN_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --quick > /dev/null 2> $O/r02z2_trace.err
tail -14 $O/r02z2_trace.err | cut -c1-160
for m in -1 0; do
TEKKEN_B200_PACK_IDS=$m timeout 900 python bench.py --workload single1g --steps 5 --no-cpu > $O/r02z2_single1g_$m.json 2> $O/r02z2_single1g_$m.err
python - $m <<'PY'
import json,sys
j=json.load(open("gpurun_out/r02z2_single1g_%s.json"%sys.argv[1])); print("single1g pack", sys.argv[1], "dev", round(j["ms_per_step"],2), "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"])
PY
done

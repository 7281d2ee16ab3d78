#!/bin/bash
# round 2, call Z: packed-id download -- all GPU tests, e2e A/B (packed 18 / 24 / off), config 3 through tk_encode
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02z_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02z_pytest.log
tail -6 $O/r02z_pytest.log
nproc > $O/r02z_nproc.txt; lscpu | head -20 >> $O/r02z_nproc.txt
run() { # name env...
  n=$1; shift
  env "$@" timeout 900 python bench.py --steps 5 --no-cpu > $O/r02z_$n.json 2> $O/r02z_$n.err
  python - "$n" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02z_%s.json"%sys.argv[1])); e=j["e2e"]
    print(sys.argv[1], "dev ms", round(j["ms_per_step"],3), "e2e ms", round(e["ms_per_step"],2), "GB/s", round(e["value"],2), "d2h", e["d2h_bytes_per_step"], "pageable", e.get("pageable_input",{}).get("ms_per_step"), "decode ms", round(j["decode"]["ms_per_step"],3))
except Exception as ex: print(sys.argv[1], "ERR", ex)
PY
}
run pack18 TEKKEN_B200_PACK_IDS=-1
run pack24 TEKKEN_B200_PACK_IDS=24
run pack0 TEKKEN_B200_PACK_IDS=0
run pack18_t6 TEKKEN_B200_PACK_IDS=-1 TEKKEN_B200_COPY_THREADS=6
TEKKEN_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --quick > /dev/null 2> $O/r02z_trace.err
timeout 900 python bench.py --workload single1g --steps 3 --no-cpu > $O/r02z_single1g.json 2> $O/r02z_single1g.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02z_single1g.json")); print("single1g dev", round(j["ms_per_step"],2), "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"])
PY

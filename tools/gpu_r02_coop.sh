#!/bin/bash
# round 2: pre-tokeniser with block-cooperative classification of non-ASCII characters (A/B), then the full GPU suite
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02coop_build.log 2>&1
run() { # name lib workload
  TEKKEN_B200_LIB=$2 TEKKEN_B200_NO_BUILD=1 timeout 600 python bench.py --workload $3 --steps 5 --no-cpu --quick --no-e2e > $O/r02coop_$1_$3.json 2> $O/r02coop_$1_$3.err
  python - "$1" "$3" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02coop_%s_%s.json"%(sys.argv[1],sys.argv[2]))); st=j["roofline"]["stage_ms"]
    print(sys.argv[1], sys.argv[2], "ms", round(j["ms_per_step"],3), "pretok", st["pretok"], "carry", st["pretok_carry"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
run old $PWD/build/variants/libtekken_b200_coop0.so mixed
run coop $PWD/tekken_rs_b200/libtekken_b200.so mixed
run old $PWD/build/variants/libtekken_b200_coop0.so english
run coop $PWD/tekken_rs_b200/libtekken_b200.so english
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02coop_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02coop_pytest.log
tail -5 $O/r02coop_pytest.log | cut -c1-300

#!/bin/bash
# round 2, call X: branch-free decode gather -- decode tests, variants, ncu launch list
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02y_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q -k "decode or roundtrip or round_trip or policy or utf8 or engine or config5 or smoke or golden" > $O/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02y_pytest.log
tail -6 $O/r02y_pytest.log
run() { # name lib
  TEKKEN_B200_LIB=$2 TEKKEN_B200_NO_BUILD=1 timeout 600 python bench.py --steps 5 --no-cpu --quick > $O/r02y_$1.json 2> $O/r02y_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02y_%s.json"%sys.argv[1])); d=j["decode"]
    print(sys.argv[1], "encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), "GB/s", round(d["value"],1), "frac", round(d["hbm_frac"],4), "exact", d["roundtrip_byte_exact"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
run base $PWD/tekken_rs_b200/libtekken_b200.so
for v in dc8_5 dc8_6; do run $v $PWD/build/variants/libtekken_b200_$v.so; done
timeout 600 python bench.py --workload english --steps 5 --no-cpu --quick > $O/r02y_english.json 2> $O/r02y_english.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02y_english.json")); d=j["decode"]
print("english encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), "GB/s", round(d["value"],1), "frac", round(d["hbm_frac"],4), d["roundtrip_byte_exact"])
PY
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"decode_gather|decode_validate" --launch-skip 2 -c 2 -o $O/r02y_decode python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > $O/r02y_ncu.log 2>&1
ls -la $O/r02y_decode.ncu-rep

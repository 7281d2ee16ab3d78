#!/bin/bash
# round 2, call V: rewritten decode kernels -- all GPU tests, then decode time for the tile-shape variants
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02v_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02v_pytest.log
tail -6 $O/r02v_pytest.log
run() { # name lib
  TEKKEN_B200_LIB=$2 TEKKEN_B200_NO_BUILD=1 timeout 600 python bench.py --steps 5 --no-cpu --quick > $O/r02v_$1.json 2> $O/r02v_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02v_%s.json"%sys.argv[1])); d=j["decode"]
    print(sys.argv[1], "encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), "GB/s", round(d["value"],1), "frac", round(d["hbm_frac"],4), "exact", d["roundtrip_byte_exact"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
run base $PWD/tekken_rs_b200/libtekken_b200.so
for v in dc8_5 dc4_8 dc4_6 dc8_3; do run $v $PWD/build/variants/libtekken_b200_$v.so; done
timeout 600 python bench.py --workload english --steps 5 --no-cpu --quick > $O/r02v_english.json 2> $O/r02v_english.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r02v_english.json")); d=j["decode"]
print("english encode ms", round(j["ms_per_step"],3), "decode ms", round(d["ms_per_step"],3), "GB/s", round(d["value"],1), "frac", round(d["hbm_frac"],4), d["roundtrip_byte_exact"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode -c 40 --csv --log-file $O/r02v_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > $O/r02v_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r02v_launches_decode.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
for r in rows[1:][-8:]: print(r[ki][:40], r[vi], r[ui])
PY

#!/bin/bash
# round 2: lane-merge with per-block key minima (A/B over the class threshold), then the parity tests on the shipped build
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { # name lib workload
  TEKKEN_B200_LIB=$2 TEKKEN_B200_NO_BUILD=1 timeout 600 python bench.py --workload $3 --steps 5 --no-cpu --quick --no-e2e > $O/r02bm_$1_$3.json 2> $O/r02bm_$1_$3.err
  python - "$1" "$3" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02bm_%s_%s.json"%(sys.argv[1],sys.argv[2]))); st=j["roofline"]["stage_ms"]
    print(sys.argv[1], sys.argv[2], "ms", round(j["ms_per_step"],3), {k[9:]:v for k,v in st.items() if k.startswith("lanemerge")})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
run off $PWD/build/variants/libtekken_b200_bm1000.so mixed
run b48 $PWD/tekken_rs_b200/libtekken_b200.so mixed
run b32 $PWD/build/variants/libtekken_b200_bm32.so mixed
run b24 $PWD/build/variants/libtekken_b200_bm24.so mixed
run off $PWD/build/variants/libtekken_b200_bm1000.so english
run b48 $PWD/tekken_rs_b200/libtekken_b200.so english
timeout 1500 python -m pytest tests -m gpu -q -x -k "fuzz or config4 or synthetic or golden or config2_mixed or piece_length or long_runs or engine_fixture" > $O/r02bm_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02bm_pytest.log
tail -4 $O/r02bm_pytest.log | cut -c1-300

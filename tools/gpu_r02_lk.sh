#!/bin/bash
# round 2: lookup kernel with the tile's stream words composed in shared memory (A/B against the shipped build)
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { # name lib
  TEKKEN_B200_LIB=$2 TEKKEN_B200_NO_BUILD=1 timeout 600 python bench.py --steps 5 --no-cpu --quick --no-e2e > $O/r02lk_$1.json 2> $O/r02lk_$1.err
  python - "$1" <<'PY'
import json,sys
try:
    j=json.load(open("gpurun_out/r02lk_%s.json"%sys.argv[1])); st=j["roofline"]["stage_ms"]
    print(sys.argv[1], "ms", round(j["ms_per_step"],3), "lookup", st["lookup"], "emit", st["emit"], "lanemerge", round(sum(v for k,v in st.items() if k.startswith("lanemerge")),3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
run base $PWD/tekken_rs_b200/libtekken_b200.so
run lkc8 $PWD/build/variants/libtekken_b200_lkc8.so
run lkc5 $PWD/build/variants/libtekken_b200_lkc5.so
TEKKEN_B200_LIB=$PWD/build/variants/libtekken_b200_lkc8.so TEKKEN_B200_NO_BUILD=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or golden or mixed or fuzz or engine" > $O/r02lk_pytest.log 2>&1; tail -3 $O/r02lk_pytest.log
TEKKEN_B200_LIB=$PWD/build/variants/libtekken_b200_lkc8.so TEKKEN_B200_NO_BUILD=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lookup_kernel --launch-skip 3 -c 1 --csv --log-file $O/r02lk_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu --quick --no-e2e > /dev/null 2>&1
grep lookup $O/r02lk_ncu.csv | cut -d, -f5,13-

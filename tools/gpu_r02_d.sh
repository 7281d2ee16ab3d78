#!/bin/bash
# round 2, call D: host-buffer engine (tests + e2e numbers) and the bucketed pair table A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02d_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02d_pytest.log
tail -30 $O/r02d_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r02d_bench_mixed.json 2> $O/r02d_bench_mixed.err
TEKKEN_B200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --quick > /dev/null 2> $O/r02d_trace.err
timeout 900 python bench.py --workload single1g --steps 3 --no-cpu > $O/r02d_bench_single1g.json 2> $O/r02d_bench_single1g.err
V=build/variants
TEKKEN_B200_NO_BUILD=1 TEKKEN_B200_LIB=$V/libtekken_b200_slot1.so timeout 300 python bench.py --steps 5 --no-cpu --quick > $O/r02d_var_slot1.json 2> $O/r02d_var_slot1.err
TEKKEN_B200_NO_BUILD=1 TEKKEN_B200_LIB=$V/libtekken_b200_b5.so TEKKEN_B200_LM_BPS=4,3,4,3,4,4,5,5,5 timeout 300 python bench.py --steps 5 --no-cpu --quick > $O/r02d_var_b5.json 2> $O/r02d_var_b5.err
TEKKEN_B200_NO_BUILD=1 TEKKEN_B200_LIB=$V/libtekken_b200_b4.so TEKKEN_B200_LM_BPS=4,3,4,3,4,4,4,4,4 timeout 300 python bench.py --steps 5 --no-cpu --quick > $O/r02d_var_b4.json 2> $O/r02d_var_b4.err
python - <<'PY'
import json
for f in ("bench_mixed","bench_single1g","var_slot1","var_b5","var_b4"):
    try:
        j=json.load(open("gpurun_out/r02d_%s.json"%f)); print(f, round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", {k:(round(v,2) if isinstance(v,float) else v) for k,v in j["e2e"].items() if k in ("value","ms_per_step","pageable_input")})
        print("   ", j["roofline"]["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 $O/r02d_bench_mixed.err $O/r02d_bench_single1g.err

#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02m_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_bounds.py -m gpu -x -q -k "config4 or synthetic or long_runs or config3_single or engine or latency or debug or golden or fuzz_batch" > $O/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02m_pytest.log
tail -6 $O/r02m_pytest.log
timeout 600 python bench.py --workload adversarial --steps 5 > $O/r02m_bench_adversarial.json 2> $O/r02m_bench_adversarial.err
TEKKEN_B200_TRACE=1 timeout 900 python bench.py --workload single1g --steps 3 --no-cpu > $O/r02m_bench_single1g.json 2> $O/r02m_bench_single1g.err
python - <<'PY'
import json
for f in ("bench_adversarial","bench_single1g"):
    try:
        j=json.load(open("gpurun_out/r02m_%s.json"%f))
        print(f, round(j["value"],3), "ms", round(j["ms_per_step"],4), "e2e", {k:v for k,v in j["e2e"].items() if k in ("value","ms_per_step")})
        print("   ", j["roofline"]["stage_ms"]); print("   cpu", j.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
grep trace $O/r02m_bench_single1g.err | tail -22

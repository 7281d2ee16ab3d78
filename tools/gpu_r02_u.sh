#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02u_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4 or synthetic or long_runs or config3_single or engine" > $O/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02u_pytest.log
tail -6 $O/r02u_pytest.log
timeout 600 python bench.py --workload adversarial --steps 5 --no-cpu > $O/r02u_bench_adversarial.json 2> $O/r02u_bench_adversarial.err
timeout 600 python bench.py --workload adversarial --pieces 7 --steps 5 --no-cpu > $O/r02u_bench_adversarial7.json 2> $O/r02u_bench_adversarial7.err
timeout 900 python bench.py --workload single1g --steps 3 --no-cpu > $O/r02u_bench_single1g.json 2> $O/r02u_bench_single1g.err
TEKKEN_B200_TRACE=1 timeout 600 python bench.py --workload single1g --steps 1 --warmup 3 --no-cpu > /dev/null 2> $O/r02u_trace_single1g.err
python - <<'PY'
import json
for f in ("bench_adversarial","bench_adversarial7","bench_single1g"):
    try:
        j=json.load(open("gpurun_out/r02u_%s.json"%f))
        print(f, round(j["value"],3), "ms", round(j["ms_per_step"],4), j.get("long_piece_stage"), j["roofline"]["stage_ms"]["longmerge"], "e2e", j["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/r02u_bench_adversarial.err

#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02l_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02l_pytest.log
tail -6 $O/r02l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --quick --latency > $O/r02l_bench_mixed.json 2> $O/r02l_bench_mixed.err
TEKKEN_B200_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --quick --latency > $O/r02l_bench_mixed_nopdl.json 2> $O/r02l_bench_mixed_nopdl.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick --docs 50000 > $O/r02l_bench_48mb.json 2> $O/r02l_bench_48mb.err
TEKKEN_B200_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick --docs 50000 > $O/r02l_bench_48mb_nopdl.json 2> $O/r02l_bench_48mb_nopdl.err
timeout 900 python bench.py --workload single1g --steps 3 --no-cpu > $O/r02l_bench_single1g.json 2> $O/r02l_bench_single1g.err
python - <<'PY'
import json
for f in ("bench_mixed","bench_mixed_nopdl","bench_48mb","bench_48mb_nopdl","bench_single1g"):
    try:
        j=json.load(open("gpurun_out/r02l_%s.json"%f))
        print(f, round(j["value"],2), "ms", round(j["ms_per_step"],4), "e2e", {k:v for k,v in j["e2e"].items() if k in ("value","ms_per_step")})
        print("   ", j["roofline"]["stage_ms"])
        if j.get("latency"): print("    lat", [(r["bytes"], round(r["gpu_us_median"],1)) for r in j["latency"]])
    except Exception as e: print(f, "ERR", e)
PY

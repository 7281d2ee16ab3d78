#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02k_build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_debug_bounds.py -m gpu -x -q > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02k_pytest.log
tail -5 $O/r02k_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --quick > $O/r02k_bench_mixed.json 2> $O/r02k_bench_mixed.err
timeout 900 python bench.py --workload single1g --steps 5 --no-cpu > $O/r02k_bench_single1g.json 2> $O/r02k_bench_single1g.err
timeout 300 python bench.py --workload english --steps 10 --no-cpu --quick > $O/r02k_bench_english.json 2> $O/r02k_bench_english.err
python - <<'PY'
import json
for f in ("bench_mixed","bench_single1g","bench_english"):
    try:
        j=json.load(open("gpurun_out/r02k_%s.json"%f))
        print(f, round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", {k:v for k,v in j["e2e"].items() if k in ("value","ms_per_step")})
        print("   ", j["roofline"]["stage_ms"]); print("    decode", j["decode"]["ms_per_step"], j["decode"]["hbm_frac"])
    except Exception as e: print(f, "ERR", e)
PY

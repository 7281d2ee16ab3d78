#!/bin/bash
# round 2, call F: config-pattern walk v2, chunk schedule, pool rounding; launch list of one step incl. decode
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02f_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_config_pattern.py tests/test_gpu_parity.py -m gpu -x -q -k "config or engine or file or latency or fuzz_batch" > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02f_pytest.log
tail -5 $O/r02f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02f_bench_mixed.json 2> $O/r02f_bench_mixed.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick --docs 50000 > $O/r02f_bench_mixed_48mb.json 2> $O/r02f_bench_mixed_48mb.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick --docs 130000 > $O/r02f_bench_mixed_128mb.json 2> $O/r02f_bench_mixed_128mb.err
timeout 600 python bench.py --split config --steps 5 --warmup 3 --quick --no-cpu > $O/r02f_bench_mixed_cfg.json 2> $O/r02f_bench_mixed_cfg.err
timeout 900 python bench.py --workload roundtrip64g --shards 4 --no-cpu > $O/r02f_bench_roundtrip4.json 2> $O/r02f_bench_roundtrip4.err
python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-e2e > $O/r02f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-e2e > $O/r02f_ncu.log 2>&1
python - <<'PY'
import json
for f in ("bench_mixed","bench_mixed_48mb","bench_mixed_128mb","bench_mixed_cfg","bench_roundtrip4"):
    try:
        j=json.load(open("gpurun_out/r02f_%s.json"%f))
        print(f, round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", {k:v for k,v in j["e2e"].items() if k in ("value","ms_per_step","seconds","pageable_input")})
        if j.get("roofline") and j["roofline"].get("stage_ms"): print("   ", j["roofline"]["stage_ms"])
        print("    decode", j.get("decode"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/r02f_bench_mixed.err $O/r02f_bench_roundtrip4.err $O/r02f_ncu.log

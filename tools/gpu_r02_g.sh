#!/bin/bash
# round 2, call G: engine with decoupled id buffers (e2e), roundtrip, config split; ncu --set full of one step + decode
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02g_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine or file or config5 or fuzz_batch or device_pointer" > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02g_pytest.log
tail -5 $O/r02g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02g_bench_mixed.json 2> $O/r02g_bench_mixed.err
TEKKEN_B200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --quick > /dev/null 2> $O/r02g_trace.err
timeout 600 python bench.py --split config --steps 5 --warmup 3 --quick --no-cpu > $O/r02g_bench_mixed_cfg.json 2> $O/r02g_bench_mixed_cfg.err
timeout 900 python bench.py --workload roundtrip64g --shards 6 --no-cpu > $O/r02g_bench_roundtrip6.json 2> $O/r02g_bench_roundtrip6.err
python - <<'PY'
import json
for f in ("bench_mixed","bench_mixed_cfg","bench_roundtrip6"):
    try:
        j=json.load(open("gpurun_out/r02g_%s.json"%f))
        print(f, round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", {k:v for k,v in j["e2e"].items() if k in ("value","ms_per_step","seconds","pageable_input")})
        if j.get("roofline") and j["roofline"].get("stage_ms"): print("   ", j["roofline"]["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-e2e"
$CMD > $O/r02g_plain.log 2>&1 || exit 1
K='regex:^(pretok_kernel|lookup_kernel|lanemerge_kernel|emit_kernel)$'
ncu --set full --clock-control none --import-source on --kernel-name "$K" --launch-skip 36 --launch-count 12 -o $O/r02g_prof_encode -f $CMD > $O/r02g_ncu_full.log 2>&1
ncu -i $O/r02g_prof_encode.ncu-rep --page raw --csv > $O/r02g_raw_encode.csv 2>/dev/null
K2='regex:^(decode_gather_kernel|decode_validate_kernel)$'
ncu --set full --clock-control none --import-source on --kernel-name "$K2" --launch-skip 2 --launch-count 2 -o $O/r02g_prof_decode -f $CMD > $O/r02g_ncu_full_decode.log 2>&1
ncu -i $O/r02g_prof_decode.ncu-rep --page raw --csv > $O/r02g_raw_decode.csv 2>/dev/null
tail -n 2 $O/r02g_ncu_full.log $O/r02g_ncu_full_decode.log | cut -c1-300

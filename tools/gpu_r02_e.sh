#!/bin/bash
# round 2, call E: everything so far (full GPU suite), bench lines, PCIe probe, launch list + full ncu of the step
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02e_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02e_pytest.log
tail -30 $O/r02e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.log 2>&1; tail -2 $O/r02e_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --latency > $O/r02e_bench_mixed.json 2> $O/r02e_bench_mixed.err
timeout 300 python tools/pcie_probe.py --gpus 1 > $O/r02e_pcie_probe_1gpu.json 2> $O/r02e_pcie_probe.err
timeout 900 python bench.py --workload roundtrip64g --shards 4 --no-cpu > $O/r02e_bench_roundtrip4.json 2> $O/r02e_bench_roundtrip4.err
python - <<'PY'
import json
for f in ("bench_mixed","bench_roundtrip4","pcie_probe_1gpu"):
    try:
        j=json.load(open("gpurun_out/r02e_%s.json"%f))
        if "by_gpus" in j: print(f, j["by_gpus"]); continue
        print(f, round(j["value"],2), "ms", round(j["ms_per_step"],3), "e2e", j["e2e"])
        if j.get("roofline") and j["roofline"].get("stage_ms"): print("   ", j["roofline"]["stage_ms"])
        print("    decode", j.get("decode")); print("    lat", j.get("latency"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/r02e_bench_mixed.err $O/r02e_bench_roundtrip4.err

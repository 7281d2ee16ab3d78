#!/bin/bash
# Developer tool: turn the outputs of tools/gpu_r02_final.sh (merged back into gpurun_out/) into the files kept under
# profiles/: bench lines, pytest log, launch list + summary, ncu --set full summary, hot lines, traffic.json, SASS evidence.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
for n in mixed reference english1m single1g adversarial roundtrip8 mixed_cfgsplit english; do cp $O/r02_bench_$n.json profiles/r02_bench_$n.json; done
cp $O/r02_pytest_gpu.txt profiles/r02_pytest_gpu.txt
cp $O/r02_launches.csv profiles/r02_launches.csv
python tools/ncu_summarize.py launches $O/r02_launches.csv > profiles/r02_launches_summary.txt
ncu -i $O/r02_prof_encode.ncu-rep --page raw --csv > $O/r02_raw_encode.csv 2>/dev/null
ncu -i $O/r02_prof_decode.ncu-rep --page raw --csv > $O/r02_raw_decode.csv 2>/dev/null
python - <<'PY'
import csv
a = list(csv.reader(open('gpurun_out/r02_raw_encode.csv')))
b = list(csv.reader(open('gpurun_out/r02_raw_decode.csv')))
assert a[0] == b[0]
w = csv.writer(open('gpurun_out/r02_raw_all.csv', 'w'))
for r in a:
    w.writerow(r)
for r in b[2:]:
    w.writerow(r)
PY
python tools/ncu_summarize.py full $O/r02_raw_all.csv $O/r02_traffic_per_kernel.json > profiles/r02_ncu_full_summary.txt
python - <<'PY'
import json
t = json.load(open('gpurun_out/r02_traffic_per_kernel.json'))
fam = {"pretok": 0, "longpiece": 0, "lookup": 0, "lanemerge": 0, "emit": 0}
for k, v in t.items():
    for f, pre in (("pretok", "pretok"), ("longpiece", "longmerge"), ("lookup", "lookup"), ("lanemerge", "lanemerge"), ("emit", "emit")):
        if k.startswith(pre):
            fam[f] += v
out = {"source": "ncu --set full, one device-resident step of `python bench.py --steps 2 --warmup 3 --no-cpu --quick --no-e2e` "
                 "(profiles/r02_ncu_full_summary.txt): dram__bytes_read.sum + dram__bytes_write.sum per launch, summed per kernel family",
       "families": fam, "step_total": sum(fam.values()), "per_kernel": {k: v for k, v in t.items() if not k.startswith('decode')},
       "decode": {k: v for k, v in t.items() if k.startswith('decode')}, "algorithmic_bytes_per_step": 2616311174}
json.dump(out, open('profiles/traffic.json', 'w'), indent=1)
print("step DRAM traffic %.2f GB" % (out["step_total"] / 1e9), fam)
PY
bash tools/sass_evidence.sh > profiles/r02_sass_evidence.txt 2>&1
: > profiles/r02_hot_lines.txt
for k in "pretok_kernel" "lookup_kernel" "emit_kernel" "lanemerge_kernel<12" "lanemerge_kernel<8"; do
  echo "==== $k (ncu --set full --import-source on; top CUDA lines by stall samples)" >> profiles/r02_hot_lines.txt
  ncu -i $O/r02_prof_encode.ncu-rep --page source --csv --print-source cuda,sass --kernel-name "regex:$k" 2>/dev/null > $O/src_k.csv
  python tools/ncu_lines.py $O/src_k.csv 14 >> profiles/r02_hot_lines.txt 2>&1
done
for k in "decode_gather" "decode_validate"; do
  echo "==== $k" >> profiles/r02_hot_lines.txt
  ncu -i $O/r02_prof_decode.ncu-rep --page source --csv --print-source cuda,sass --kernel-name "regex:$k" 2>/dev/null > $O/src_k.csv
  python tools/ncu_lines.py $O/src_k.csv 14 >> profiles/r02_hot_lines.txt 2>&1
done
tail -3 profiles/r02_launches_summary.txt

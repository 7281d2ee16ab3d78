#!/bin/bash
# round 2, multi-GPU call (N = number of visible GPUs): PCIe probe at 1..N, one-call-all-GPUs test, torchrun bench at N (and N/2),
# the config-5 round trip dealt over N ranks
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $O/r02p${N}_smi.txt 2>&1; nvidia-smi topo -m >> $O/r02p${N}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/r02p${N}_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_visible_gpus" > $O/r02p${N}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p${N}_pytest.log
tail -3 $O/r02p${N}_pytest.log
timeout 300 python tools/pcie_probe.py > $O/r02p${N}_pcie_probe.json 2> $O/r02p${N}_pcie_probe.err; head -c 1500 $O/r02p${N}_pcie_probe.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > $O/r02p${N}_bench_n$N.json 2> $O/r02p${N}_bench_n$N.err
H=$((N/2))
if [ $H -ge 2 ] && [ $H -ne 2 ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $H --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $H --steps 5 --warmup 3 --no-cpu --quick > $O/r02p${N}_bench_n$H.json 2> $O/r02p${N}_bench_n$H.err
fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload roundtrip64g --shards 64 --no-cpu > $O/r02p${N}_bench_roundtrip64.json 2> $O/r02p${N}_bench_roundtrip64.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02p*_bench_*.json")):
    try:
        j=json.load(open(f)); print(f.split("/")[-1], "n", j["n_gpus"], "value", round(j["value"],2), j["unit"], "ms", round(j["ms_per_step"],3), "e2e", j.get("e2e",{}).get("value"), j.get("e2e_one_call_all_gpus",{}).get("by_gpus"))
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 $O/r02p${N}_bench_n$N.err

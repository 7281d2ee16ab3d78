#!/bin/bash
# round 2, 2-GPU call: the one-call-all-GPUs path on real devices, torchrun bench at N=2, PCIe probe at 1-2 GPUs
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r02p2_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/r02p2_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_visible_gpus or engine" > $O/r02p2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p2_pytest.log
tail -4 $O/r02p2_pytest.log
timeout 300 python tools/pcie_probe.py > $O/r02p2_pcie_probe.json 2> $O/r02p2_pcie_probe.err; cat $O/r02p2_pcie_probe.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02p2_bench_n2.json 2> $O/r02p2_bench_n2.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r02p2_bench_n2.json")); print("n2", round(j["value"],2), j["e2e"]["value"], j.get("e2e_one_call_all_gpus"))
except Exception as e: print("ERR", e)
PY
tail -n 5 $O/r02p2_bench_n2.err

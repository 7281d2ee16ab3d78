#!/bin/bash
# tools/gpurun_retry_n.sh N SCRIPT LOG [TIMEOUT] -- like gpurun_retry.sh, on N GPUs of one box
N=$1; S=$2; L=$3; T=${4:-1500}
for i in $(seq 1 40); do
  while ! python -c "import sys; from tekken_rs_b200 import _build; sys.exit(1 if _build.needs_build() else 0)"; do sleep 20; done
  /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "bash $S" > $L 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $L; then sleep 180; continue; fi
  break
done
tail -5 $L

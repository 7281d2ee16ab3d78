#!/bin/bash
# tools/gpurun_retry.sh SCRIPT LOG [TIMEOUT] -- run `bash SCRIPT` on a GPU box, retrying while the pool answers "transient"
# (nothing charged).  Waits while the in-tree library is older than its sources (an edit is in progress).
S=$1; L=$2; T=${3:-2400}
for i in $(seq 1 40); do
  while ! python -c "import sys; from tekken_rs_b200 import _build; sys.exit(1 if _build.needs_build() else 0)"; do sleep 20; done
  /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > $L 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $L; then sleep 120; continue; fi
  break
done
tail -5 $L

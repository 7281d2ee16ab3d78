#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02t2_build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/r02t2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02t2_pytest.log
tail -30 $O/r02t2_pytest.log | cut -c1-300

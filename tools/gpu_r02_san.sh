#!/bin/bash
# round 2: compute-sanitizer memcheck over the tests of the kernels that changed last (decode, latency paths, long pieces)
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02san_build.log 2>&1
K="decode_token_lengths or single_sequence_decode or single_text_latency or decode_policies or decode_errors or config4_adversarial or golden or ragged or fuzz_window"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/r02san_plain.log 2>&1 || { tail -5 $O/r02san_plain.log; echo "plain run failed"; exit 1; }
tail -2 $O/r02san_plain.log
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/r02san_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/r02san_pytest.log 2>&1; echo "sanitizer rc=$?"
tail -3 $O/r02san_pytest.log | cut -c1-200
tail -5 $O/r02san_memcheck.log | cut -c1-200
grep -c "Invalid\|out of bounds\|misaligned" $O/r02san_memcheck.log

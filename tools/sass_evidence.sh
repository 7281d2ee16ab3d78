#!/bin/bash
# Evidence that the kernels use Blackwell/Hopper-era hardware paths (no GPU needed): SASS mnemonics of the built library.
#   UBLKCP            cp.async.bulk (TMA 1-D bulk copy) -- tile staging of pretok_kernel / lookup_kernel
#   SYNCS.*           mbarrier (expect_tx arrive, try_wait) the bulk copy completes on
#   ACQBULK           griddepcontrol.wait -- programmatic dependent launch along the encode chain
# Usage: tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
cd "$(dirname "$0")/.."
LIB=tekken_rs_b200/libtekken_b200.so
echo "# cuobjdump -sass $LIB (sm_100a) -- $(date -u +%Y-%m-%d)"
echo "# instruction counts over the whole library"
cuobjdump -sass $LIB | grep -oE "UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|ACQBULK|UTMALDG|HMMA|UTC[A-Z]*MMA|LDGSTS[.A-Z0-9]*" | sort | uniq -c
echo "# functions that contain them"
cuobjdump -sass $LIB | awk '/Function :/{f=$3} /UBLKCP|SYNCS|ACQBULK/{split($0,a," "); k=f" "; for(i=1;i<=NF;i++) if ($i ~ /^(UBLKCP|SYNCS|ACQBULK)/) {print f, $i}}' | sort | uniq -c | c++filt | sed 's/(.*//' | awk '{print $1, $2, $3, $4}' | sort -k2 | uniq
echo "# the bulk copy and its mbarrier in lookup_kernel"
cuobjdump -sass $LIB | awk '/Function : .*lookup_kernel/{p=1} /Function :/&&!/lookup_kernel/{p=0} p&&/UBLKCP|SYNCS|ACQBULK/{print}'
echo "# registers / shared memory per kernel (cuobjdump -res-usage)"
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/ Function /\n/;s/Fatbin.*//' | c++filt | awk '{$1=$1};1' | cut -c1-220

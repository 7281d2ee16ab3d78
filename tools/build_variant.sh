#!/bin/bash
# build an experimental variant of the library next to the real one: tools/build_variant.sh NAME "-DFOO=1 -DBAR=2"
# -> build/variants/libtekken_b200_NAME.so (use with TEKKEN_B200_LIB=... TEKKEN_B200_NO_BUILD=1)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
S=tekken_rs_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $2 \
  -o build/variants/libtekken_b200_$1.so $S/tk_kernels.cu $S/tk_decode.cu $S/tk_api.cu $S/tk_host.cpp
echo built build/variants/libtekken_b200_$1.so

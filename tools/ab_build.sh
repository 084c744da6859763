#!/bin/bash
# builds kernel variants for A/B runs: tools/ab_build.sh name "-DFOO=1 -DBAR=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -shared -Xcompiler -fPIC $flags \
     -o build/ab/libeuc_$name.so euc_b200/csrc/euc_b200.cu -ldl -lrt &
done
wait
ls -la build/ab

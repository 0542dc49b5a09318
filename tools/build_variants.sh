#!/bin/bash
# development: build A/B variants of libsvi_ls.so into svinet_b200/lib/variants/ (args: name "-Dflags" ...)
set -e
cd "$(dirname "$0")/.."
mkdir -p svinet_b200/lib/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC,-O2 -Xptxas -O3 $flags -Iinclude \
     -shared -o svinet_b200/lib/variants/libsvi_ls_$name.so svinet_b200/csrc/svi_ls.cu svinet_b200/csrc/svi_fa2.cu &
done
wait

#!/bin/bash
# build experimental variants of the library into build/variants/ (git-ignored; travels with gpurun)
#   scripts/build_variants.sh name "-DFLAG=1 -DOTHER=2" [name2 "flags2" ...]
set -e
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
       -o build/variants/libaesmc_$name.so aesmc_b200/csrc/*.cu &
done
wait
ls -la build/variants/

#!/bin/bash
# Builds theboss_b200/lib/libbossperm_<NAME>.so: the production objects with minors_kernel.cu recompiled under extra flags
# (compile-time tuning knobs of K3).  Used with BOSSPERM_LIB=... for A/B measurements; never shipped.
#   scripts/build_variant.sh NAME [-DK3_...=...]...
set -e
cd "$(dirname "$0")/../theboss_b200/csrc"
NAME=$1; shift
mkdir -p build/var_$NAME
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a"
$NV -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 "$@" -Xptxas -v -c minors_kernel.cu -o build/var_$NAME/minors_kernel.o 2> build/var_$NAME/minors_kernel.ptxas.log
$NV -shared -cudart static -ccbin /usr/bin/g++ -o ../lib/libbossperm_$NAME.so build/api.o build/glynn_kernel.o build/util_kernels.o build/guan_kernel.o build/sampler_kernel.o build/var_$NAME/minors_kernel.o
echo "built ../lib/libbossperm_$NAME.so"

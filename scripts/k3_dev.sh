#!/bin/bash
# Compile ONE k3_minors_kernel instantiation to a scratch object and print registers, spills, the static schedule of its
# loops and the register-read model of the term loop.
#   scripts/k3_dev.sh LPG C [THREADS] [extra nvcc flags, e.g. -DK3_MINB3_MAX_C=12]
set -e
cd "$(dirname "$0")/../theboss_b200/csrc"
LPG=$1; C=$2; T=${3:-128}; shift; shift; shift || true
mkdir -p build/dev
OUT=build/dev/k3_${LPG}_${C}_${T}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2 \
  -DK3_DEV_LPG=$LPG -DK3_DEV_C=$C -DK3_DEV_THREADS=$T "$@" -Xptxas -v -c minors_kernel.cu -o $OUT 2>&1 | grep -A2 "k3_minors_kernel" | grep -E "registers|spill"
python ../../scripts/sass_sched.py $OUT "k3_minors_kernelILi${LPG}ELi${C}ELi${T}E" 2 -loops | tail -12
python ../../scripts/sass_rf.py $OUT "k3_minors_kernelILi${LPG}ELi${C}ELi${T}E"

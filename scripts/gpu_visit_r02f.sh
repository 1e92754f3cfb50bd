#!/bin/bash
# Round 2, visit: rewritten K2 (uniform loop, tree product, host prep for small batches) -- parity, A/B against the previous build.
set -x
mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
OUT=gpurun_out/k2_ab.txt
: > $OUT
echo "== previous build" >> $OUT
BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_old.so timeout 120 python scripts/time_k2_single.py >> $OUT 2>&1
echo "== this build" >> $OUT
timeout 120 python scripts/time_k2_single.py >> $OUT 2>&1
cat $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2_perm_kernel -c 1 -o gpurun_out/k2_c2_r02 python scripts/profile_k2.py 10000 > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/ncu_k2.log

#!/bin/bash
# Round 2, visit 2: A/B of compile-time K3 variants (scripts/build_variant.sh -> lib/libbossperm_<name>.so) and block-sizing knobs.
#   gpurun --timeout 900 -- 'bash scripts/gpu_visit_r02c.sh'
set -x
mkdir -p gpurun_out
OUT=gpurun_out/ab_k3_c.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $OUT 2>&1
AB_TAG=default timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
for v in regrow12 minb3_12 minb3_12r minb3_10; do
  BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_$v.so AB_TAG=$v timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
done
for tpg in 1024 4096 8192; do
  BP_K3_TPG=$tpg AB_TAG=default timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
  BP_K3_TPG=$tpg BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_regrow12.so AB_TAG=regrow12 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
done
for cap in 8 32 64; do
  BP_K3_CAP=$cap AB_TAG=default timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
done
cat $OUT

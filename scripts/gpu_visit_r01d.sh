#!/bin/bash
# GPU visit: A/B of the wide K3 variants with the inner row re-read from shared memory (x8) and 3 blocks/SM (m10, m12).
mkdir -p gpurun_out
L=theboss_b200/lib
cp $L/libbossperm.so $L/keep.so
{
  timeout 100 python scripts/ab_k3.py 3 short
  for v in x8_m8 x8_m10 x8_m12; do
    cp $L/libbossperm_$v.so $L/libbossperm.so
    AB_TAG=$v timeout 100 python scripts/ab_k3.py 3 short
    AB_TAG=$v timeout 100 python scripts/profile_c5.py 32
  done
  cp $L/keep.so $L/libbossperm.so
} > gpurun_out/ab_k3_d.txt 2>&1
cat gpurun_out/ab_k3_d.txt

#!/bin/bash
# GPU visit: <*,8> variants with the inner row re-read from shared memory (no spills at 3 blocks/SM) vs registers.
mkdir -p gpurun_out
L=theboss_b200/lib
cp $L/libbossperm.so $L/keep.so
{
  timeout 100 python scripts/profile_c5.py 32
  timeout 100 python scripts/profile_c5.py 32
  timeout 100 python scripts/ab_k3.py 3 short
  cp $L/libbossperm_x7_m8.so $L/libbossperm.so
  AB_TAG=x7 timeout 100 python scripts/profile_c5.py 32
  AB_TAG=x7 timeout 100 python scripts/profile_c5.py 32
  AB_TAG=x7 timeout 100 python scripts/ab_k3.py 3 short
  cp $L/keep.so $L/libbossperm.so
} > gpurun_out/ab_k3_e.txt 2>&1
cat gpurun_out/ab_k3_e.txt

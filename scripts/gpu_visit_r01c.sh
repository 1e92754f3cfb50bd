#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# GPU visit: A/B of K3 block shapes for the large steps (256-thread blocks, 3 blocks/SM for C = 9, 10).
mkdir -p gpurun_out
L=theboss_b200/lib
{
  timeout 100 python scripts/ab_k3.py 3 short
  BP_K3_WIDE_MIN_K=21 timeout 100 python scripts/ab_k3.py 3 short
  BP_K3_BIG_THREADS=1 timeout 100 python scripts/ab_k3.py 3 short
  BP_K3_BIG_THREADS=1 BP_K3_TPG=1024 timeout 100 python scripts/ab_k3.py 3 short
  cp $L/libbossperm.so $L/keep.so; cp $L/libbossperm_minb3.so $L/libbossperm.so
  AB_TAG=minb3 timeout 100 python scripts/ab_k3.py 3 short
  cp $L/keep.so $L/libbossperm.so
  timeout 100 python scripts/profile_c5.py 32
  BP_K3_BIG_THREADS=1 timeout 100 python scripts/profile_c5.py 32
} > gpurun_out/ab_k3_c.txt 2>&1
cat gpurun_out/ab_k3_c.txt
timeout 300 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3

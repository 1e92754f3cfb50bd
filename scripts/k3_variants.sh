#!/bin/bash
# Times the GCC-B sampling run (n = 24, 4096 samples) and the dilated n = 30 run under the K3 column-split variants.
mkdir -p gpurun_out
{
for mc in 0 8; do
  echo "== BP_K3_MAX_C=$mc"
  BP_K3_MAX_C=$mc timeout 300 python scripts/profile_k3.py 24 4096 2 2>&1 | grep "samples/s"
done
for mc in 0 7 12; do
  echo "== c5 BP_K3_MAX_C=$mc"
  BP_K3_MAX_C=$mc timeout 300 python scripts/profile_c5.py 32 2>&1 | grep "samples/s"
done
} | tee gpurun_out/k3_variants.txt

#!/bin/bash
# Times the GCC-B sampling run (n = 24, 4096 samples) under the K3 column-split variants (BP_K3_MAX_C).
mkdir -p gpurun_out
for mc in 0 8 7; do
  echo "== BP_K3_MAX_C=$mc"
  BP_K3_MAX_C=$mc timeout 300 python scripts/profile_k3.py 24 4096 3 2>&1 | grep "samples/s"
done | tee gpurun_out/k3_variants.txt

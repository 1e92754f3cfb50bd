#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# Times GCC-B sampling runs under different K3 work-sizing knobs (BP_K3_TPG = terms per lane group and block,
# BP_K3_CAP = chunk blocks per sample when samples are plentiful).
mkdir -p gpurun_out
{
for cfg in "192 3" "1024 8" "1024 32" "2048 8" "2048 16" "2048 32" "2048 64" "4096 32" "8192 32"; do
  set -- $cfg
  echo "== BP_K3_TPG=$1 BP_K3_CAP=$2"
  export BP_K3_TPG=$1 BP_K3_CAP=$2
  timeout 300 python scripts/profile_k3.py 24 4096 2 2>&1 | grep "samples/s" | tail -1
  timeout 300 python scripts/profile_k3.py 24 1024 2 2>&1 | grep "samples/s" | tail -1
  timeout 300 python scripts/profile_k3.py 20 16384 2 2>&1 | grep "samples/s" | tail -1
  timeout 300 python scripts/profile_c5.py 32 2>&1 | grep "samples/s"
done
} | tee gpurun_out/k3_variants.txt

#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# Round 2, visit 4: engine 2 with three blocks per SM up to C = 12 (lib variant u3) against the plain engine 2 build.
set -x
mkdir -p gpurun_out
OUT=gpurun_out/ab_k3_e.txt
: > $OUT
export BP_K3_ENGINE=2 BP_K3_TREE_MAX_C=17
AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 >> $OUT 2>&1
BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_u3.so AB_TAG=eng2_u3 timeout 120 python scripts/ab_k3.py 3 >> $OUT 2>&1
BP_K3_TREE_MAX_C=12 BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_u3.so AB_TAG=eng2_u3 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
BP_K3_TREE_MAX_C=9 BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_u3.so AB_TAG=eng2_u3 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
BP_K3_TREE_MAX_C=6 BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_u3.so AB_TAG=eng2_u3 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
cat $OUT
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
BOSSPERM_LIB=$PWD/theboss_b200/lib/libbossperm_u3.so timeout 300 ncu --metrics $M --clock-control none -k regex:k3u?_minors --csv --log-file gpurun_out/k3_steps_eng2_u3.csv python scripts/profile_k3.py 24 4096 0 > gpurun_out/k3_steps_eng2_u3.log 2>&1

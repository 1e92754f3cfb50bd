#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# Round 2, visit 3: engine 2 of K3 (one uniform term loop): parity under BP_K3_ENGINE=2, A/B against engine 1, per-step pipe.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_visit_r02d.sh'
set -x
mkdir -p gpurun_out
export BP_K3_TREE_MAX_C=17
BP_K3_ENGINE=2 timeout 900 python -m pytest tests/test_gpu_minors.py tests/test_gpu_sampling.py tests/test_gpu_properties.py tests/test_gpu_edges.py tests/test_gpu_bobs.py tests/test_gpu_dispatch.py tests/test_gpu_zz_reference_runs.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_eng2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_eng2.log
tail -30 gpurun_out/pytest_gpu_eng2.log
OUT=gpurun_out/ab_k3_d.txt
: > $OUT
BP_K3_ENGINE=1 AB_TAG=eng1 timeout 120 python scripts/ab_k3.py 3 >> $OUT 2>&1
BP_K3_ENGINE=2 AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 >> $OUT 2>&1
BP_K3_ENGINE=2 BP_K3_WARP_MAX_K=0 AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 >> $OUT 2>&1
BP_K3_ENGINE=2 BP_K3_TREE_MAX_C=12 AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
BP_K3_ENGINE=2 BP_K3_TREE_MAX_C=8 AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
BP_K3_ENGINE=2 BP_K3_TREE_MAX_C=6 AB_TAG=eng2 timeout 120 python scripts/ab_k3.py 3 short >> $OUT 2>&1
cat $OUT
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
BP_K3_ENGINE=2 timeout 300 ncu --metrics $M --clock-control none -k regex:k3u?_minors --csv --log-file gpurun_out/k3_steps_eng2.csv python scripts/profile_k3.py 24 4096 0 > gpurun_out/k3_steps_eng2.log 2>&1
BP_K3_ENGINE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3u_minors_kernel -s 22 -c 1 -o gpurun_out/k3_n24_eng2 python scripts/profile_k3.py 24 2048 0 > gpurun_out/k3_n24_eng2.log 2>&1
ls -la gpurun_out | tail -5

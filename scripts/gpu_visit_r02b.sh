#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# Round 2, visit 1: parity of the tree engine (K3) + A/B against the scan engine + per-step FP64 pipe + flop cross-check.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_visit_r02b.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
# A/B: engines and lane limits (one process per setting)
: > gpurun_out/ab_k3.txt
BP_K3_ENGINE=0 AB_TAG=scan timeout 120 python scripts/ab_k3.py 3 >> gpurun_out/ab_k3.txt 2>&1
AB_TAG=tree timeout 120 python scripts/ab_k3.py 3 >> gpurun_out/ab_k3.txt 2>&1
for c in 16 12 8 6; do BP_K3_TREE_MAX_C=$c AB_TAG=tree timeout 120 python scripts/ab_k3.py 3 short >> gpurun_out/ab_k3.txt 2>&1; done
BP_K3_WARP_MAX_K=12 AB_TAG=tree timeout 120 python scripts/ab_k3.py 3 short >> gpurun_out/ab_k3.txt 2>&1
BP_K3_WARP_MAX_K=0 AB_TAG=tree timeout 120 python scripts/ab_k3.py 3 short >> gpurun_out/ab_k3.txt 2>&1
cat gpurun_out/ab_k3.txt
# per-step time and FP64 pipe of a whole n = 24 run, tree engine (default) and two lanes from k = 13
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 300 ncu --metrics $M --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_tree.csv python scripts/profile_k3.py 24 4096 0 > gpurun_out/k3_steps_tree.log 2>&1
BP_K3_TREE_MAX_C=12 timeout 300 ncu --metrics $M --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_tree_c12.csv python scripts/profile_k3.py 24 4096 0 > gpurun_out/k3_steps_tree_c12.log 2>&1
# flop cross-check: executed FP64 thread instructions over every k3_minors launch of one 512-sample run vs the algorithmic count
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_fp64_counts_n24.csv python scripts/profile_k3.py 24 512 0 > gpurun_out/k3_fp64_counts_n24.log 2>&1
tail -2 gpurun_out/k3_fp64_counts_n24.log
# full capture of the k = 24 step (launch index 22 of the k3_minors launches) of a 2048-sample run
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_minors_kernel -s 22 -c 1 -o gpurun_out/k3_n24_tree python scripts/profile_k3.py 24 2048 0 > gpurun_out/k3_n24_tree.log 2>&1
ls -la gpurun_out

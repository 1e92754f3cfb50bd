#!/bin/bash
# Round 2 profiling visit: evidence for profiles/ (no benchmark values are taken under the profiler).
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
# (1) flop cross-check: executed FP64 thread instructions over every k3_minors launch of one 512-sample n = 24 run
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_fp64_counts_n24.csv python scripts/profile_k3.py 24 512 0 > gpurun_out/k3_fp64_counts_n24.log 2>&1
tail -2 gpurun_out/k3_fp64_counts_n24.log
# (2) the FP64 peak probe itself: pipe utilisation of the roofline denominator
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.avg.per_second,launch__registers_per_thread \
    --clock-control none -k regex:fp64_peak_kernel --csv --log-file gpurun_out/fp64_peak_probe.csv python -c "
import sys; sys.path.insert(0, '.')
from theboss_b200 import _native
print(_native.default_handle(0).fp64_peak(100.0))" > gpurun_out/fp64_peak_probe.log 2>&1
tail -1 gpurun_out/fp64_peak_probe.log
# (3) launch list of a uniform-loss run (C5(i), 4096 samples)
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/c5i_launches.csv python scripts/profile_c5i.py 4096 0 > gpurun_out/c5i_launches.log 2>&1
# (4) launch list of a short bench run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
# (5) the dilated config 5(ii) steps: per-step pipe of a 32-sample run
timeout 600 ncu --metrics $M --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/c5ii_steps.csv python scripts/profile_c5.py 32 > gpurun_out/c5ii_steps.log 2>&1
ls -la gpurun_out | tail -8

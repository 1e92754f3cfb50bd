#!/bin/bash
# HISTORICAL: this visit ran on an earlier build; some BP_K3_* knobs it sets (ENGINE, WIDE_MIN_K, MAX_C) were removed with the engines they selected.
# GPU visit for the one-warp K3 blocks: tests, smoke, A/B timings of the dispatch knobs, sanitizer, bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
{
  timeout 200 python scripts/ab_k3.py 3
  BP_K3_WARP_MAX_K=0 timeout 200 python scripts/ab_k3.py 3
  BP_K3_WARP_MAX_K=12 timeout 200 python scripts/ab_k3.py 3
  BP_K3_WARP_MAX_K=8 timeout 200 python scripts/ab_k3.py 3
  BP_K3_WIDE_MIN_K=17 timeout 200 python scripts/ab_k3.py 3
  BP_K3_WIDE_MIN_K=19 timeout 200 python scripts/ab_k3.py 3
} > gpurun_out/ab_k3.txt 2>&1
cat gpurun_out/ab_k3.txt
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_n24.csv python scripts/profile_k3.py 24 4096 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_n20.csv python scripts/profile_k3.py 20 16384 0 > /dev/null 2>&1
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitizer_$tool.log
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
cat gpurun_out/bench_reference.json

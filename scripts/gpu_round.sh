#!/bin/bash
# One GPU visit: tests, smoke, bench, launch list, one full ncu capture of each kernel (K1, K2, K3).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:glynn_block4_kernel -s 1 -c 1 -f -o gpurun_out/k1_n30 python scripts/profile_k1.py 30 > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/ncu_k1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_minors_kernel -s 22 -c 1 -f -o gpurun_out/k3_n24 python scripts/profile_k3.py 24 2048 0 > gpurun_out/ncu_k3.log 2>&1
tail -2 gpurun_out/ncu_k3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_perm_kernel -s 1 -c 1 -f -o gpurun_out/k2_c2 python scripts/profile_k2.py 10000 > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k3_launches.csv python scripts/profile_k3.py 24 4096 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_n24.csv python scripts/profile_k3.py 24 4096 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_steps_n20.csv python scripts/profile_k3.py 20 16384 0 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/c5i_launches.csv python scripts/profile_c5i.py 4096 0 > /dev/null 2>&1
bash scripts/time_sampling.sh
timeout 100 python scripts/profile_c5i.py 4096 3 | tee -a gpurun_out/time_sampling.txt
timeout 100 python scripts/ab_k3.py 3 | tee -a gpurun_out/time_sampling.txt
timeout 100 python scripts/time_c1.py | tee -a gpurun_out/time_sampling.txt
timeout 100 python scripts/c3_latency.py | tee -a gpurun_out/time_sampling.txt

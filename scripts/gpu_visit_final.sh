#!/bin/bash
# Final GPU visit of the round: tests, smoke, bench (both arms), launch list of the bench command, K1 full capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:glynn_block4_kernel -s 1 -c 1 -f -o gpurun_out/k1_n30 python scripts/profile_k1.py 30 > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_k1.log
python scripts/k1_sweep.py 20 22 23 24 26 28 30 31 32 34 35 36 38 40 > gpurun_out/k1_sweep.txt 2>&1
python scripts/time_shard.py >> gpurun_out/k1_sweep.txt 2>&1
tail -6 gpurun_out/k1_sweep.txt

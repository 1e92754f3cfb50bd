#!/bin/bash
# First GPU visit of round 2: everything that was added after the last GPU visit of round 1 gets its first measurement.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_visit_r02a.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
# 1. parity: the whole GPU suite (the module test_gpu_zz_reference_runs.py is new: reference-run fixtures of rows f1 / f3,
#    config 2 / 3 outputs of the reference at n = 20 / k <= 16, seeded GCC-B runs at n = 12 .. 16)
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
# 2. bench line with the new gcc_sampling.roofline object, reference arm
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("K1 roofline:", {k: d["roofline"][k] for k in ("achieved", "peak", "frac")})
print("sampling:", d["gcc_sampling"]["value"], "samples/s;", d["gcc_sampling"].get("roofline"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
# 3. cross-check of the sampling flop accounting (ncu's dadd + dmul + 2 dfma within ~10 % of the algorithmic count):
#    FP64 thread-instruction counters over every k3_minors launch of ONE n = 24 run of 512 samples, next to the count
#    scripts/profile_k3.py prints for the same run
timeout 600 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k3_minors --csv --log-file gpurun_out/k3_fp64_counts_n24.csv python scripts/profile_k3.py 24 512 0 > gpurun_out/k3_fp64_counts_n24.log 2>&1
tail -2 gpurun_out/k3_fp64_counts_n24.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/k3_fp64_counts_n24.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Metric Name" in r)
col = {k: i for i, k in enumerate(rows[hdr])}
tot = {}
for r in rows[hdr + 1:]:
    name = r[col["Metric Name"]]
    tot[name] = tot.get(name, 0.0) + float(r[col["Metric Value"]].replace(",", ""))
g = lambda op: tot.get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum", 0.0)
print(f"executed FP64 flops over all k3_minors launches: {g('dadd') + g('dmul') + 2 * g('dfma'):.4e}  (dadd {g('dadd'):.3e} dmul {g('dmul'):.3e} dfma {g('dfma'):.3e})")
PY
# 4. launch list of the bench (share of K1 in the step) and of a BOBS request (slices)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1

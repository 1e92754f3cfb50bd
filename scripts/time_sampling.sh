#!/bin/bash
# Standard timing set of the sampling path: GCC-B n=24 (4096 and 1024 samples), n=20 (16384), dilated n=30 (32),
# config 2 (K2) and the single n=24 step (config 3).
mkdir -p gpurun_out
{
timeout 300 python scripts/profile_k3.py 24 4096 2 2>&1 | grep "samples/s" | tail -1
timeout 300 python scripts/profile_k3.py 24 1024 2 2>&1 | grep "samples/s" | tail -1
timeout 300 python scripts/profile_k3.py 20 16384 2 2>&1 | grep "samples/s" | tail -1
timeout 300 python scripts/profile_c5.py 32 2>&1 | grep "samples/s"
timeout 300 python scripts/profile_k2.py 10000 2>&1 | tail -1
timeout 300 python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from tests import workloads
from theboss_b200 import _native
h = _native.default_handle(0)
for cf in (False, True):
    U, s, t = workloads.c3_step(24, 48, cf)
    h.gccb_pmf(U, s, t)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); h.gccb_pmf(U, s, t); ts.append(time.perf_counter() - t0)
    print(f"c3 step n=24 collision_free={cf}: best {min(ts)*1e6:.1f} us, median {sorted(ts)[10]*1e6:.1f} us")
PY
} | tee gpurun_out/time_sampling.txt

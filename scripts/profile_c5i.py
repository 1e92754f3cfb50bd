"""Timing / ncu driver for BASELINE config 5 (i): uniform losses eta = 0.5, n = 30, m = 60, through bp_gccb_simulate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = _native.default_handle(0)
U, _, s = workloads.c5_lossy(30, 60)
for r in range(reps + 1):
    t0 = time.perf_counter()
    out = h.gccb_simulate(U, s, S, eta=0.5, seed=5)
    dt = time.perf_counter() - t0
    print(f"c5(i) eta=0.5 n=30 m=60 S={S}: {dt*1e3:.2f} ms, {S/dt:.1f} samples/s, mean detected {out.sum(axis=1).mean():.2f}, launches so far {h.launch_count()}", flush=True)

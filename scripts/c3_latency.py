"""Latency of one GCC-B step through the host-pointer API (bp_gccb_pmf) at n = 24, m = 48: wall clock per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native
h = _native.default_handle(0)
for cf in (False, True):
    U, s, t = workloads.c3_step(24, 48, cf)
    for _ in range(20): h.gccb_pmf(U, s, t)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); h.gccb_pmf(U, s, t); ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"c3 step n=24 collision_free={cf}: best {ts[0]*1e6:.1f} us, median {ts[100]*1e6:.1f} us, p90 {ts[180]*1e6:.1f} us")

"""Cost of ONE permanent through the batched multiplicity kernel (what RyserPermanentCalculator.compute_permanent() pays):
device time (CUDA events around the host-pointer call) and wall clock, best of 50, plus the kernel launches per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

h = _native.default_handle(0)
for n, m in ((4, 8), (8, 16), (14, 28), (20, 40)):
    U = workloads.haar(m, n)
    rng = np.random.RandomState(n)
    S = np.zeros((1, m), dtype=np.uint8); T = np.zeros((1, m), dtype=np.uint8)
    for j in rng.randint(0, m, n): S[0, j] += 1
    for j in rng.randint(0, m, n): T[0, j] += 1
    h.perm_batched(U, S, T)
    dev, wall = [], []
    l0 = h.launch_count()
    for _ in range(50):
        t0 = time.perf_counter()
        h.timer_start()
        out = h.perm_batched(U, S, T)
        dev.append(h.timer_stop())
        wall.append(time.perf_counter() - t0)
    launches = (h.launch_count() - l0) / 50
    print(f"n={n} m={m}: device {min(dev) * 1e3:7.1f} us  wall {min(wall) * 1e6:7.1f} us  launches/call {launches:.0f}  perm {out[0]:.6e}", flush=True)
for n, m, B in ((20, 40, 10000), (12, 24, 10000), (30, 60, 64)):
    U, S, T = workloads.c2_batch(n, m, B)
    h.perm_batched(U, S, T)
    dev = []
    for _ in range(3):
        h.timer_start(); out = h.perm_batched(U, S, T); dev.append(h.timer_stop())
    print(f"batch n={n} m={m} B={B}: device {min(dev):8.3f} ms  checksum {np.abs(out).sum():.12e}", flush=True)

"""Timing driver for K2: BASELINE config 2 (10^4 batched n=20 permanents with repeated rows/columns)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

items = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
h = _native.default_handle(0)
U, S, T = workloads.c2_batch(items=items)
terms = d_sum = 0.0
for b in range(items):
    cs, ct = np.prod(S[b].astype(np.float64) + 1), np.prod(T[b].astype(np.float64) + 1)
    walk_s = cs <= ct
    d = np.count_nonzero(T[b] if walk_s else S[b])
    terms += min(cs, ct) / 2
    d_sum += min(cs, ct) / 2 * (2 * d + 6 * 19 + 4)
for rep in range(3):
    t0 = time.perf_counter()
    out = h.perm_batched(U, S, T)
    dt = time.perf_counter() - t0
    print(f"items={items}: {dt*1e3:.2f} ms, {items/dt:.0f} permanents/s, {d_sum/dt/1e12:.2f} TFLOP/s useful (SURVEY 8d formula), terms {terms:.3e}", flush=True)

"""A/B timing of the lossy / few-sample sampling workloads (late steps: a handful of samples spread over many chunk blocks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native
h = _native.default_handle(0)
tag = os.environ.get("AB_TAG", "")
def run(name, U, s, S, eta=-1.0, reps=5):
    best = 1e9
    for _ in range(reps + 1):
        t0 = time.perf_counter(); out = h.gccb_simulate(U, s, S, eta=eta, seed=5); best = min(best, time.perf_counter() - t0)
    chk = int((out.astype(np.int64) * np.arange(1, out.shape[1] + 1)).sum())
    print(f"[{tag}] {name}: {best * 1e3:9.3f} ms  checksum {chk}", flush=True)
U, _, s = workloads.c5_lossy(30, 60)
run("c5(i) S=4096", U, s, 4096, eta=0.5)
run("c5(i) S=10000", U, s, 10000, eta=0.5)
run("c5(i) S=512", U, s, 512, eta=0.5)
for n, S in ((20, 512), (20, 64), (24, 64), (16, 256)):
    run(f"gccb n={n} S={S}", workloads.haar(2 * n, n), np.array([1] * n + [0] * n, dtype=np.int32), S, reps=3)

"""A/B timing of the GCC-B sampling loop under the K3 dispatch knobs read from the environment
(BP_K3_WARP_MAX_K, BP_K3_WARP2_MIN_K, BP_K3_TREE_MAX_C, BP_K3_TPG, BP_K3_CAP).  One process per setting (the knobs are read once);
prints the best of `reps` wall-clock runs per workload plus a checksum of the samples (all settings must agree)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
h = _native.default_handle(0)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("BP_K3_")) or "defaults"


def run(name, U, s, S, eta=-1.0):
    best, out = 1e9, None
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        out = h.gccb_simulate(U, s, S, eta=eta, seed=5)
        best = min(best, time.perf_counter() - t0)
    chk = int((out.astype(np.int64) * np.arange(1, out.shape[1] + 1)).sum())
    print(f"[{tag}] {name}: {best * 1e3:9.3f} ms  {S / best:12.1f} samples/s  checksum {chk}", flush=True)


short = len(sys.argv) > 2 and sys.argv[2] == "short"
tag = (os.environ.get("AB_TAG", "") + " " + tag).strip()
for n, S in (((24, 4096), (20, 16384)) if short else ((24, 4096), (20, 16384), (16, 16384), (12, 16384), (8, 16384), (20, 512))):
    U = workloads.haar(2 * n, n)
    s = np.array([1] * n + [0] * n, dtype=np.int32)
    run(f"gccb n={n} m={2 * n} S={S}", U, s, S)
U, _, s = workloads.c5_lossy(30, 60)
if not short:
    run("c5(i) uniform eta=0.5 n=30 m=60 S=4096", U, s, 4096, eta=0.5)

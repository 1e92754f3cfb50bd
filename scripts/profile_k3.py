"""Driver for ncu / timing of the GCC-B sampling loop: S samples at n photons in m = 2n modes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
h = _native.default_handle(0)
U = workloads.haar(2 * n, n)
s = np.array([1] * n + [0] * n, dtype=np.int32)
for r in range(reps + 1):
    t0 = time.perf_counter()
    out = h.gccb_simulate(U, s, S, seed=5)
    dt = time.perf_counter() - t0
    print(f"n={n} S={S}: {dt*1e3:.2f} ms, {S/dt:.1f} samples/s, launches so far {h.launch_count()}", flush=True)
# algorithmic flops of the run: per step T * (22k - 36), T = prod(t_j + 1) / 2 over the outputs sampled so far
tot = 0.0
for row in out:
    # order of arrival inside a sample is not returned; bound by the final occupation: sum over prefixes is <= 2x the last step
    t = row.astype(np.float64)
    tot += np.prod(t + 1) / 2
print(f"sum over samples of the final-step walk length bound prod(t+1)/2 (with all n outputs): {tot:.3e}")

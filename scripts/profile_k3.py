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
# algorithmic flops of the run (sum over samples and steps of ceil(prod(t_j + 1) / 2) * (22 k - 36), counted on the run's own
# outputs): compare with ncu's dadd + dmul + 2 * dfma over all k3_minors launches of ONE run (reps = 0)
import bench
flops = bench.sampling_algorithmic_flops(out)
print(f"algorithmic flops of one run: {flops:.4e} ({flops / S:.4e} per sample; collision-free worst case "
      f"{sum(2.0 ** (k - 2) * (22 * k - 36) for k in range(2, n + 1)):.4e}); useful TFLOP/s at the last timing: {flops / dt / 1e12:.2f}")

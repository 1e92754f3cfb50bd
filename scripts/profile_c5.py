"""Timing driver for BASELINE config 5 (ii): GCC-B on the 2m-mode dilation of a lossy n=30, m=60 network."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import workloads
from theboss_b200 import _native
from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = _native.default_handle(0)
U, U_lossy, s = workloads.c5_lossy(n, 2 * n)
big = np.ascontiguousarray(prepare_interferometer_matrix_in_expanded_space(U_lossy))
s_big = np.concatenate([s, np.zeros(2 * n, dtype=np.int32)])
h.gccb_simulate(big, s_big, 4, seed=1)
t0 = time.perf_counter()
res = h.gccb_simulate(big, s_big, S, seed=5)
dt = time.perf_counter() - t0
print(f"c5(ii) n={n} m={2*n} (dilated {4*n}) S={S}: {dt:.3f} s, {S/dt:.2f} samples/s, checksum {int((res * np.arange(1, res.shape[1] + 1)).sum())}", flush=True)

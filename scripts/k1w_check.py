import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from tests import workloads
from theboss_b200 import _native
from oracle import pyoracle as orc
h = _native.default_handle(0)
for N in (35, 36, 39, 40):
    A = workloads.c4_matrix(N)
    for lo, hi in ((0, 1 << 18), (1 << 20, (1 << 20) + (1 << 17)), (64, (1 << 17) + 64 * 3), ((1 << 30) + 4096, (1 << 30) + 4096 + (1 << 16))):
        p = h.glynn_matrix_range(A, lo, hi)
        got = complex(p[0] + p[1], p[2] + p[3])
        want = orc.glynn_range(A, lo, hi, "ld")
        print(N, lo, hi, abs(got - want) / max(abs(want), 1e-300), flush=True)
rng = np.random.RandomState(36)
N = 36
d = np.exp(1j * rng.uniform(0, 2 * np.pi, N)) * rng.uniform(0.8, 1.2, N)
A = np.zeros((N, N), dtype=np.complex128); A[rng.permutation(N), np.arange(N)] = d
got = h.glynn_matrix(A); print("perm36", abs(got - np.prod(d)) / abs(np.prod(d)))

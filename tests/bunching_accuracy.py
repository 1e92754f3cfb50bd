"""Relative error of K2 / K3 under heavy bunching (the cases of tests/test_gpu_edges.py) against the 80-bit oracle, next to the error
of the double-precision restatement of the reference (which uses pow() for repeated factors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
from tests import workloads
from theboss_b200 import _native
from oracle import pyoracle as orc

h = _native.default_handle(0)
U = workloads.haar(8, 40)
S = np.array([[20, 20, 0, 0, 0, 0, 0, 0], [40, 0, 0, 0, 0, 0, 0, 0], [10, 10, 10, 10, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 7, 33]], dtype=np.uint8)
T = np.array([[10, 10, 10, 10, 0, 0, 0, 0], [0, 0, 40, 0, 0, 0, 0, 0], [5, 5, 5, 5, 5, 5, 5, 5], [13, 0, 0, 27, 0, 0, 0, 0]], dtype=np.uint8)
got = h.perm_batched(U, S, T)
for b in range(len(S)):
    want = orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "ld")
    ref = orc.guan_permanent(U, S[b], T[b], orc.CHIN_HUH, "d")
    print(f"K2 item {b}: kernel rel err {abs(got[b] - want) / abs(want):.3e}   reference-in-double rel err {abs(ref - want) / abs(want):.3e}")
U = workloads.haar(6, 41)
s = np.array([10, 9, 8, 7, 7, 0], dtype=np.int32)
t = np.array([0, 20, 0, 20, 0, 0], dtype=np.int32)
got = h.minors(U, s, t)
want = orc.submatrices(U, s, t, orc.CHIN_HUH, "ld")
ref = orc.submatrices(U, s, t, orc.CHIN_HUH, "d")
print(f"K3 k=41: kernel rel err {np.abs(got - want).max() / np.abs(want).max():.3e}   reference-in-double rel err {np.abs(ref - want).max() / np.abs(want).max():.3e}")

"""Tiny driver for ncu: two n=N Glynn permanents through the C ABI (first = warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import workloads
from theboss_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
h = _native.default_handle(0)
A = workloads.c4_matrix(n)
for _ in range(2):
    print(h.glynn_matrix(A))

"""Tiny driver for ncu: two range partials of an N x N Glynn permanent (N >= 35: glynn_pair4_kernel) over 2^LOG aligned Gray steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import workloads
from theboss_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
log = int(sys.argv[2]) if len(sys.argv) > 2 else 26
h = _native.default_handle(0)
A = workloads.c4_matrix(n)
for _ in range(2):
    print(h.glynn_matrix_range(A, 0, 1 << log))

import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from tests import workloads
from theboss_b200 import _native
h = _native.default_handle(0)
U, s, t = workloads.c3_step(24, 48, True)
for _ in range(5): h.gccb_pmf(U, s, t)

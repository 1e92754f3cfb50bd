"""Warp-pair K1 kernel alone, for compute-sanitizer --tool synccheck / racecheck (scripts/sanitize.sh covers it inside the full job)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import workloads
from theboss_b200 import _native
h = _native.default_handle(0)
for N, lo, hi in ((36, 0, 1 << 17), (39, 64, (1 << 16) + 64), (40, 96, (1 << 16) + 160)):
    print("K1 wide", N, h.glynn_matrix_range(workloads.c4_matrix(N), lo, hi)[:2], flush=True)

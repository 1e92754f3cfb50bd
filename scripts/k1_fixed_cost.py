"""Fixed cost of one K1 call at n = 30: device-resident ranges of 2^17 .. 2^26 Gray steps, CUDA events (min of 20)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads
from theboss_b200 import _native
h = _native.Handle(0, stream_ptr=torch.cuda.current_stream(0).cuda_stream)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
A = workloads.c4_matrix(N)
dA = torch.from_numpy(A.view(np.float64).copy()).cuda()
out = torch.zeros(4, dtype=torch.float64, device="cuda")
resident = len(sys.argv) > 2 and sys.argv[2] == "resident"
if resident:
    h.glynn_set_resident(dA.data_ptr())      # the constant-bank image is written once, not per call
    print("matrix declared resident")
for lg in (17, 19, 21, 22, 23, 24, 26):
    if lg > N - 1: break
    hi = 1 << lg
    for _ in range(3):
        h.glynn_matrix_range_dev(dA.data_ptr(), N, 0, hi, out.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        h.timer_start(); h.glynn_matrix_range_dev(dA.data_ptr(), N, 0, hi, out.data_ptr()); ts.append(h.timer_stop())
    print(f"N={N} 2^{lg} steps: {min(ts)*1e3:8.1f} us")
h.timer_start(); ms = h.timer_stop()
print(f"empty timer pair: {ms*1e3:.1f} us")

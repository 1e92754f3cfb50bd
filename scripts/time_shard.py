"""Times kernel K1 on one 1/G slice of the n=30 Gray range (what each rank runs at G GPUs), device-resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads
from theboss_b200 import _native
h = _native.Handle(0, stream_ptr=torch.cuda.current_stream(0).cuda_stream)
A = workloads.c4_matrix(30)
dA = torch.from_numpy(A.view(np.float64).copy()).cuda()
out = torch.zeros(4, dtype=torch.float64, device="cuda")
for G in (1, 2, 4, 8):
    hi = (1 << 29) // G
    for _ in range(3):
        h.glynn_matrix_range_dev(dA.data_ptr(), 30, 0, hi, out.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        h.timer_start(); h.glynn_matrix_range_dev(dA.data_ptr(), 30, 0, hi, out.data_ptr()); ts.append(h.timer_stop())
    ms = min(ts)
    print(f"G={G}: {ms:.4f} ms per shard -> {1e3/ms:.1f} permanents/s if perfectly parallel, {236 * hi / (ms * 1e-3) / 1e12:.2f} TFLOP/s useful per GPU")

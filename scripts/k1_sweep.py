"""K1 throughput over N: 2^28 aligned Gray steps of a Haar N x N submatrix, device-resident, CUDA events.
BP_K1_BULK_MAX_N limits the block-4 bulk kernel (tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads
from theboss_b200 import _native
h = _native.Handle(0, stream_ptr=torch.cuda.current_stream(0).cuda_stream)
peak = h.fp64_peak(100.0)
out = torch.zeros(4, dtype=torch.float64, device="cuda")
Ns = [int(x) for x in sys.argv[1:]] or [20, 22, 23, 24, 26, 28, 30, 31, 32, 34, 35, 36, 38, 40]
for N in Ns:
    A = workloads.c4_matrix(N)
    dA = torch.from_numpy(np.ascontiguousarray(A).view(np.float64).copy()).cuda()
    steps = min(1 << 28, 1 << (N - 1))
    for _ in range(2):
        h.glynn_matrix_range_dev(dA.data_ptr(), N, 0, steps, out.data_ptr())
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        h.timer_start(); h.glynn_matrix_range_dev(dA.data_ptr(), N, 0, steps, out.data_ptr()); ts.append(h.timer_stop())
    ms = min(ts)
    tf = (8 * N - 4) * steps / (ms * 1e-3) / 1e12
    slots = (6 * N - 4) * steps * 2 / (ms * 1e-3) / 1e12
    print(f"N={N:2d}: {ms:8.3f} ms for 2^{int(np.log2(steps))} steps, {tf:6.2f} TFLOP/s useful = {tf/peak:.3f} of peak {peak:.2f}, FP64 issue slots {slots/peak:.3f}", flush=True)

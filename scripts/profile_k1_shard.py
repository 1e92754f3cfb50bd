"""ncu driver: one 1/8 shard (2^26 Gray steps) of the n = 30 permanent through the resident + exchange-free range entry point."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import workloads
from theboss_b200 import _native
h = _native.Handle(0, stream_ptr=torch.cuda.current_stream(0).cuda_stream)
A = workloads.c4_matrix(30)
dA = torch.from_numpy(A.view(np.float64).copy()).cuda()
out = torch.zeros(4, dtype=torch.float64, device="cuda")
h.glynn_set_resident(dA.data_ptr())
for _ in range(3):
    h.glynn_matrix_range_dev(dA.data_ptr(), 30, 0, 1 << 26, out.data_ptr())
torch.cuda.synchronize()
print(out.cpu().numpy())

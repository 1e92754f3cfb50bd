"""Peer-memory partial exchange of the sharded Glynn permanent (bp_exchange_*, bp_glynn_matrix_range_exchange)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("N", [24, 36])
def test_exchange_with_a_single_rank_returns_the_range_partial(N):
    """(N = 24: block-4 bulk kernel; N = 36: the warp-pair kernel -- the exchange runs in whichever kernel finishes the sum)"""
    import torch
    from theboss_b200 import _native
    h = _native.Handle(0)
    A = workloads.c4_matrix(N)
    dA = torch.from_numpy(np.ascontiguousarray(A).view(np.float64).reshape(-1).copy()).cuda()
    d_part = torch.zeros(4, dtype=torch.float64, device="cuda")
    d_all = torch.zeros(4, dtype=torch.float64, device="cuda")
    ipc = h.exchange_create(1, 0)
    assert len(ipc) == 64
    h.exchange_connect([ipc])
    for lo, hi in ((0, 1 << min(N - 1, 24)), (64, 1 << 20), (100, 5000), (7, 7)):      # bulk kernel, bulk + tails, generic, empty
        h.glynn_matrix_range_dev(dA.data_ptr(), N, lo, hi, d_part.data_ptr())
        h.glynn_matrix_range_exchange(dA.data_ptr(), N, lo, hi, d_all.data_ptr())
        h.synchronize()
        assert torch.equal(d_part.cpu(), d_all.cpu()), (lo, hi)
    # host-buffer form of the collective: one call, same bits
    for lo, hi in ((0, 1 << 20), (100, 5000)):
        want = h.glynn_matrix_range(A, lo, hi)
        got = h.glynn_matrix_range_exchange_host(A, lo, hi, 1)
        assert got.shape == (1, 4) and tuple(got[0]) == tuple(want)
    h.exchange_destroy()
    with pytest.raises(_native.BossPermError):
        h.glynn_matrix_range_exchange(dA.data_ptr(), N, 0, 64, d_all.data_ptr())
    with pytest.raises(_native.BossPermError):
        h.glynn_matrix_range_exchange_host(A, 0, 64, 1)
    h.close()


def test_resident_matrix_reuses_and_refreshes_the_constant_bank_image():
    import torch
    from theboss_b200 import _native
    h, other = _native.Handle(0), _native.Handle(0)
    N = 25
    A1, A2 = workloads.c4_matrix(N), workloads.c4_matrix(N)[::-1].copy()
    want1, want2 = h.glynn_matrix(A1), h.glynn_matrix(A2)
    assert want1 != want2
    dA = torch.from_numpy(np.ascontiguousarray(A1).view(np.float64).reshape(-1).copy()).cuda()
    d_part = torch.zeros(4, dtype=torch.float64, device="cuda")
    scale = 2.0 ** -(N - 1)

    def perm():
        h.glynn_matrix_range_dev(dA.data_ptr(), N, 0, 1 << (N - 1), d_part.data_ptr())
        h.synchronize()
        p = d_part.cpu().numpy()
        return complex((p[0] + p[1]) * scale, (p[2] + p[3]) * scale)

    h.glynn_set_resident(dA.data_ptr())
    assert perm() == want1 and perm() == want1                      # second call: no copy into the constant bank
    assert other.glynn_matrix(A2) == want2                          # another handle overwrites the image ...
    assert perm() == want1                                          # ... and the resident one notices
    dA.copy_(torch.from_numpy(np.ascontiguousarray(A2).view(np.float64).reshape(-1).copy()))
    h.glynn_set_resident(dA.data_ptr())                             # same pointer, new contents
    assert perm() == want2 and perm() == want2
    h.glynn_set_resident(None)
    assert perm() == want2
    h.close(); other.close()


_WORKER = r"""
import os, sys
sys.path.insert(0, %(repo)r)
import numpy as np, torch, torch.distributed as dist
from tests import workloads
from theboss_b200.distributed import ShardedGlynnPermanent
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(0)
job = ShardedGlynnPermanent(24, device=0, exchange=sys.argv[4])
out = []
for seed in (24, 25, 26, 27, 28):                       # five calls: both halves of the slot buffers, in turn
    A = workloads.haar(48, seed)[:24, :24].copy()
    out.append(job.compute(A))
print("RESULT", rank, job.exchange, " ".join(repr(x) for x in out), flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_exchange_through_peer_memory(tmp_path):
    """Two processes (ranks) on the one visible GPU map each other's slot buffers through CUDA IPC and run the sharded permanent:
    identical results on both ranks, equal to the single-GPU value, and equal to the NCCL-free "nccl" fallback path's sum order."""
    from theboss_b200 import _native
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"repo": REPO})
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", port, "peer"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    lines = [next(l for l in o[0].splitlines() if l.startswith("RESULT")) for o in outs]
    res = [l.split(" ", 3) for l in lines]
    assert res[0][2] == res[1][2] == "peer", lines
    assert res[0][3] == res[1][3], lines                              # bit-identical on both ranks
    h = _native.default_handle(0)
    got = [complex(x) for x in res[0][3].split()]
    for seed, g in zip((24, 25, 26, 27, 28), got):
        want = h.glynn_matrix(workloads.haar(48, seed)[:24, :24].copy())
        assert abs(g - want) <= 1e-13 * abs(want)

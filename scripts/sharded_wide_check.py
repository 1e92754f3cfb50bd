"""torchrun check: the sharded Glynn permanent at N = 36 (warp-pair kernel + in-kernel peer exchange) on a scaled permutation matrix,
whose permanent is the product of the scales; every rank must return the same bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from theboss_b200.distributed import ShardedGlynnPermanent

dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 36
rng = np.random.RandomState(N)
d = np.exp(1j * rng.uniform(0, 2 * np.pi, N)) * rng.uniform(0.8, 1.2, N)
A = np.zeros((N, N), dtype=np.complex128)
A[rng.permutation(N), np.arange(N)] = d
job = ShardedGlynnPermanent(N, device=torch.cuda.current_device())
got = [job.compute(A) for _ in range(3)]
want = np.prod(d)
vals = [None] * world
dist.all_gather_object(vals, repr(got))
if rank == 0:
    print("exchange:", job.exchange, "rel err", abs(got[0] - want) / abs(want), "identical on all ranks and calls:", len(set(vals)) == 1 and len(set(got)) == 1, flush=True)
    assert abs(got[0] - want) <= 1e-10 * abs(want) and len(set(vals)) == 1 and len(set(got)) == 1
dist.barrier()
dist.destroy_process_group()

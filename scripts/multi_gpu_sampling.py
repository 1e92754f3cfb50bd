"""Sharded GCC-B sampling over the ranks of a torchrun job (NCCL): BASELINE config 3/5 style runs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_sampling.py [n] [samples]

Every rank draws its contiguous slice of the samples (Philox keyed by the global sample index), one final
gather follows; rank 0 also runs the whole job alone and checks that the sharded result is bit-identical."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]     # before torch is imported: keeps NCCL's banner off stdout
import numpy as np
import torch
import torch.distributed as dist

from tests import workloads
from theboss_b200 import _native
from theboss_b200.distributed import sharded_gccb_simulate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
S = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
U = workloads.haar(2 * n, n)
s = np.array([1] * n + [0] * n, dtype=np.int32)
sharded_gccb_simulate(U, s, S, seed=1, device=local)            # warm-up at full size (scratch allocation, NCCL init)
if rank == 0:
    _native.default_handle(local).gccb_simulate(U, s, S, seed=1)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = sharded_gccb_simulate(U, s, S, seed=5, device=local)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    t1 = time.perf_counter()
    alone = _native.default_handle(local).gccb_simulate(U, s, S, seed=5)
    dt1 = time.perf_counter() - t1
    print(json.dumps({"n": n, "m": 2 * n, "samples": S, "n_gpus": world, "seconds": dt, "samples_per_s": S / dt,
                      "single_gpu_seconds": dt1, "single_gpu_samples_per_s": S / dt1,
                      "identical_to_single_gpu_run": bool(np.array_equal(out, alone))}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

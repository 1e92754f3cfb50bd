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
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"          # plain | uniform | nonuniform (BASELINE config 5 i / ii)
check_alone = (sys.argv[4] != "nocheck") if len(sys.argv) > 4 else True
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eta = -1.0
if mode == "plain":
    U = workloads.haar(2 * n, n)
    s = np.array([1] * n + [0] * n, dtype=np.int32)
else:
    U, U_lossy, s = workloads.c5_lossy(n, 2 * n)
    if mode == "uniform":
        eta = 0.5
    else:
        from theboss_b200.boson_sampling_utilities.boson_sampling_utilities import prepare_interferometer_matrix_in_expanded_space
        U = np.ascontiguousarray(prepare_interferometer_matrix_in_expanded_space(U_lossy))
        s = np.concatenate([s, np.zeros(2 * n, dtype=np.int32)])
_simulate = sharded_gccb_simulate
sharded_gccb_simulate = lambda U_, s_, S_, seed, device: _simulate(U_, s_, S_, eta=eta, seed=seed, device=device)
sharded_gccb_simulate(U, s, S if mode != "nonuniform" else min(S, 64 * world), seed=1, device=local)   # warm-up (scratch allocation, NCCL init)
if rank == 0 and check_alone:
    _native.default_handle(local).gccb_simulate(U, s, S, eta=eta, seed=1)
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
    line = {"mode": mode, "n": n, "m": 2 * n, "modes_simulated": int(U.shape[0]), "samples": S, "n_gpus": world, "seconds": dt,
            "samples_per_s": S / dt, "particles_conserved": bool((out.sum(axis=1) <= n).all())}
    if check_alone:
        t1 = time.perf_counter()
        alone = _native.default_handle(local).gccb_simulate(U, s, S, eta=eta, seed=5)
        dt1 = time.perf_counter() - t1
        line.update({"single_gpu_seconds": dt1, "single_gpu_samples_per_s": S / dt1,
                     "identical_to_single_gpu_run": bool(np.array_equal(out, alone))})
    print(json.dumps(line))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

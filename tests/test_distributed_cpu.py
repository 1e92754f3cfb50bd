"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: Gray-range sharding + all-gather of
double-double partials + fixed-order combine, and the final sample gather.  The per-rank compute is
stood in for by the CPU oracle (test infrastructure); on the GPU box the same code path runs kernel K1
over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import workloads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from theboss_b200.distributed import allgather_partials, combine_partials, gather_samples, gray_shard, shard_bounds

        A = workloads.c4_matrix(N)
        lo, hi = gray_shard(N, world, rank)
        part = orc.glynn_range(A, lo, hi, "ld")
        mine = torch.tensor([part.real, 0.0, part.imag, 0.0], dtype=torch.float64)
        allp = allgather_partials(mine)
        value = combine_partials(allp.numpy(), N)
        # sample shards of unequal size: rank r owns [lo, hi) of 7 samples, 5 modes
        s_lo, s_hi = shard_bounds(7, world, rank)
        local = torch.arange(s_lo * 5, s_hi * 5, dtype=torch.int32).reshape(s_hi - s_lo, 5)
        allsamples = gather_samples(local)
        out_q.put((rank, value, allp.numpy().copy(), allsamples.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_permanent_and_sample_gather_world2():
    from oracle import pyoracle as orc
    N, world = 13, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    [p.start() for p in procs]
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    truth = orc.glynn_matrix(workloads.c4_matrix(N), "ld")
    values = [r[1] for r in results]
    assert values[0] == values[1]                       # fixed-order sum: identical bits on every rank
    assert abs(values[0] - truth) <= 1e-14 * abs(truth)
    assert np.array_equal(results[0][2], results[1][2])
    want = np.arange(35, dtype=np.int32).reshape(7, 5)
    assert np.array_equal(results[0][3], want) and np.array_equal(results[1][3], want)


def _dynamic_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import handle_standin
        from theboss_b200 import _native
        from theboss_b200.distributed import dynamic_gccb_simulate
        standin = handle_standin.OracleHandle()
        _native.default_handle = lambda device=0: standin       # CPU stand-in of the device handle (test infrastructure)
        U = workloads.haar(6, 3)
        s = np.array([1, 1, 1, 0, 0, 0], dtype=np.int32)
        got = dynamic_gccb_simulate(U, s, 37, seed=9, device=0, batch=5)
        again = dynamic_gccb_simulate(U, s, 37, seed=9, device=0, batch=8)       # a second call uses a fresh counter
        out_q.put((rank, got.copy(), again.copy(), standin.gccb_simulate(U, s, 37, seed=9)))
    finally:
        dist.destroy_process_group()


def test_samples_handed_out_on_demand_equal_the_single_rank_run_world2():
    """dynamic_gccb_simulate (batches of samples drawn from a shared counter in the process group's store): whatever rank draws
    which batch, every rank ends with the array a single rank would have produced -- the generator is keyed by the global
    sample index."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dynamic_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    results = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for _, got, again, single in results:
        assert got.shape == (37, 6) and np.all(got.sum(axis=1) == 3)
        assert np.array_equal(got, single) and np.array_equal(again, single)

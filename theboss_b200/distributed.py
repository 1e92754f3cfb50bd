"""Multi-GPU sharding of the permanent hot path (one process per GPU, torch.distributed / NCCL).

The reference has no distributed backend; what shards is

* one large Glynn permanent: the 2^(N-1) Gray steps are split into ``world_size`` contiguous
  slices, every rank runs kernel K1 on its slice and produces ONE double-double complex partial
  (32 bytes); a single all-gather follows and every rank adds the partials in rank order, so the
  result is identical on all ranks and independent of NCCL's algorithm choice;
* batches of independent samples / items: contiguous slices per rank, no traffic until the final
  gather (``shard_bounds``).

torch is used for plumbing only (device buffers, streams, the collective).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``total`` units owned by ``rank`` (sizes differ by at most 1)."""
    q, r = divmod(int(total), int(world_size))
    lo = q * rank + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gray_shard(N: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Slice of the 2^(N-1) Gray steps of an N x N Glynn permanent owned by ``rank``."""
    return shard_bounds(1 << (N - 1), world_size, rank)


def combine_partials(partials: np.ndarray, N: int) -> complex:
    """Fixed-order (rank order) sum of double-double partials {re_hi, re_lo, im_hi, im_lo}, scaled by
    2^-(N-1) (glynn_gray_permanent_calculator.py:69 in the reference).  Python floats: TwoSum in
    plain IEEE double arithmetic, so every rank gets the same bits."""
    def two_sum(a, b):
        s = a + b
        bb = s - a
        return s, (a - (s - bb)) + (b - bb)

    acc = [[0.0, 0.0], [0.0, 0.0]]
    for p in np.asarray(partials, dtype=np.float64).reshape(-1, 4):
        for c in range(2):
            s, e = two_sum(acc[c][0], float(p[2 * c]))
            e += acc[c][1] + float(p[2 * c + 1])
            hi = s + e
            acc[c] = [hi, e - (hi - s)]
    scale = 2.0 ** (-(N - 1))
    return complex((acc[0][0] + acc[0][1]) * scale, (acc[1][0] + acc[1][1]) * scale)


def allgather_partials(partial, group=None):
    """All-gather one 4-double partial per rank (NCCL on GPU tensors, gloo on CPU tensors) and return
    the (world_size, 4) tensor in rank order.  Single-process runs return the partial itself."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial.reshape(1, 4).clone()
    world = dist.get_world_size(group)
    out = torch.empty(4 * world, dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out.reshape(world, 4)


def gather_samples(local_samples, group=None):
    """Final gather of a sharded sampling run: (S_local, m) integer tensors -> (S_total, m) in rank order.
    The only collective of a sampling job; shards may differ in size by one."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_samples
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_samples.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_samples.shape[0]], dtype=torch.int64, device=local_samples.device), group=group)
    cap = int(max(int(x.item()) for x in sizes))
    padded = torch.zeros((cap, local_samples.shape[1]), dtype=local_samples.dtype, device=local_samples.device)
    padded[: local_samples.shape[0]] = local_samples
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[: int(n.item())] for p, n in zip(parts, sizes)], dim=0)


class ShardedGlynnPermanent:
    """perm(A) of one explicit N x N matrix over all ranks of a process group.

    ``compute(A)`` takes a HOST matrix (the user-facing call: upload, K1 on this rank's Gray slice,
    all-gather of 32-byte partials, fixed-order sum); ``enqueue_resident()`` / ``finish()`` is the same
    with the matrix already resident in HBM (bench.py's device-timed leg).  ``exchange``: "peer" (default; partials travel
    through peer memory inside the K1 kernel, falling back to NCCL when the GPUs cannot map each other) or "nccl"."""

    def __init__(self, N: int, device: Optional[int] = None, group=None, exchange: str = "peer"):
        import torch
        import torch.distributed as dist

        from . import _native

        self._torch, self._dist, self.group = torch, dist, group
        self.N = int(N)
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.lo, self.hi = gray_shard(self.N, self.world_size, self.rank)
        stream = torch.cuda.current_stream(self.device)
        self._stream_ptr = stream.cuda_stream
        self.handle = _native.Handle(self.device, stream_ptr=stream.cuda_stream)
        dev = torch.device("cuda", self.device)
        self.d_A = torch.zeros(self.N * self.N * 2, dtype=torch.float64, device=dev)
        self.d_part = torch.zeros(4, dtype=torch.float64, device=dev)
        self.d_all = torch.zeros(4 * self.world_size, dtype=torch.float64, device=dev)
        self.h_A = torch.zeros(self.N * self.N * 2, dtype=torch.float64).pin_memory()
        self.h_all = torch.zeros(4 * self.world_size, dtype=torch.float64).pin_memory()
        self.handle.glynn_set_resident(self.d_A.data_ptr())
        self.exchange = "none" if self.world_size == 1 else "nccl"
        if self.world_size > 1 and exchange != "nccl":
            self._connect_peer_exchange()

    def _connect_peer_exchange(self) -> None:
        """Partial exchange over peer memory (NVLink): the last block of K1 stores this rank's 32 bytes into every peer's slot
        buffer and waits for theirs -- no NCCL launch per permanent.  Every rank must agree: if any rank cannot map a peer
        (no peer access between the GPUs, IPC unavailable) all of them keep the NCCL all-gather."""
        torch, dist = self._torch, self._dist
        ok, handles = 1, None
        try:
            mine = self.handle.exchange_create(self.world_size, self.rank)
            handles = [None] * self.world_size
            dist.all_gather_object(handles, mine, group=self.group)
            self.handle.exchange_connect(handles)
        except Exception:   # noqa: BLE001 -- fall back together
            ok = 0
        on_gpu = dist.get_backend(self.group) != "gloo"      # (gloo: the two-process single-GPU test)
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", self.device) if on_gpu else torch.device("cpu"))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.exchange = "peer"
        else:
            try:
                self.handle.exchange_destroy()
            except Exception:   # noqa: BLE001
                pass

    def upload(self, A: np.ndarray) -> None:
        A = np.ascontiguousarray(A, dtype=np.complex128)
        assert A.shape == (self.N, self.N)
        self.h_A.numpy()[:] = A.view(np.float64).reshape(-1)
        self.d_A.copy_(self.h_A, non_blocking=True)
        self.handle.glynn_set_resident(self.d_A.data_ptr())     # new contents: the kernel's constant-bank image is rebuilt once

    def enqueue_resident(self) -> None:
        """K1 on this rank's slice + the partial exchange, all on the current stream, no host sync."""
        assert self._torch.cuda.current_stream(self.device).cuda_stream == self._stream_ptr, (
            "ShardedGlynnPermanent is bound to the torch stream that was current when it was built")
        if self.exchange == "peer":
            self.handle.glynn_matrix_range_exchange(self.d_A.data_ptr(), self.N, self.lo, self.hi, self.d_all.data_ptr())
            return
        self.handle.glynn_matrix_range_dev(self.d_A.data_ptr(), self.N, self.lo, self.hi, self.d_part.data_ptr())
        if self.world_size > 1:
            self._dist.all_gather_into_tensor(self.d_all, self.d_part, group=self.group)
        else:
            self.d_all.copy_(self.d_part)

    def finish(self) -> complex:
        self.h_all.copy_(self.d_all, non_blocking=True)
        self._torch.cuda.current_stream(self.device).synchronize()
        return combine_partials(self.h_all.numpy(), self.N)

    def compute(self, A: np.ndarray) -> complex:
        """One permanent with HOST buffers in and out.  With the peer-memory exchange (or a single rank) this is ONE C-ABI call:
        staged upload, this rank's slice of K1, the in-kernel exchange, download of all partials."""
        if self.exchange == "peer":
            return combine_partials(self.handle.glynn_matrix_range_exchange_host(A, self.lo, self.hi, self.world_size), self.N)
        if self.world_size == 1:
            assert np.shape(A) == (self.N, self.N)
            return combine_partials(np.asarray(self.handle.glynn_matrix_range(A, self.lo, self.hi), dtype=np.float64).reshape(1, 4), self.N)
        self.upload(A)
        self.enqueue_resident()
        return self.finish()

    @property
    def h2d_bytes(self) -> int:
        return self.N * self.N * 16

    @property
    def d2h_bytes(self) -> int:
        return 32 * self.world_size


def sharded_gccb_simulate(U, input_state, n_samples: int, eta: float = -1.0, seed: int = 0, device: Optional[int] = None,
                          group=None, gather: bool = True):
    """GCC-B sampling run split over the ranks of a process group: rank r draws the contiguous slice
    ``shard_bounds(n_samples, world, r)`` of the samples with the counter-based generator keyed by the GLOBAL
    sample index (``first_sample``), so the concatenation equals the single-GPU run bit for bit.  No traffic
    until the final gather.  Returns an (n_samples, m) int32 array on every rank
    (or only this rank's slice with ``gather=False``)."""
    import torch
    import torch.distributed as dist

    from . import _native

    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    dev = torch.cuda.current_device() if device is None else int(device)
    lo, hi = shard_bounds(n_samples, world, rank)
    Um = _native.as_matrix(U)
    s = _native.as_state(input_state, Um.shape[0])
    local = _native.default_handle(dev).gccb_simulate(Um, s, hi - lo, eta=eta, seed=seed, first_sample=lo)
    if not gather or world == 1:
        return local
    t = torch.from_numpy(local).to(torch.device("cuda", dev))
    return gather_samples(t, group).cpu().numpy()


_dynamic_calls = 0


def dynamic_gccb_simulate(U, input_state, n_samples: int, eta: float = -1.0, seed: int = 0, device: Optional[int] = None,
                          group=None, batch: int = 64, timer=None):
    """GCC-B sampling run whose samples are handed out to the ranks in batches of ``batch`` on demand (a shared counter in the
    process group's store), for workloads whose per-sample cost varies by orders of magnitude -- the dilated lossy networks of
    BASELINE config 5(ii), where a sample costs between ~1e6 and 3e11 flops depending on the outcome it happens to draw.  The
    generator is keyed by the GLOBAL sample index, so the result is the same array as the single-GPU run whatever rank drew
    which batch.  Returns the full (n_samples, m) int32 array on every rank (one all-reduce of the scattered rows at the end).
    ``timer``: optional list that receives this rank's summed device time of its bp_gccb_simulate calls in milliseconds."""
    import torch
    import torch.distributed as dist

    from . import _native

    global _dynamic_calls
    _dynamic_calls += 1
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    dev = torch.cuda.current_device() if device is None else int(device)
    Um = _native.as_matrix(U)
    s = _native.as_state(input_state, Um.shape[0])
    h = _native.default_handle(dev)
    out = np.zeros((int(n_samples), Um.shape[0]), dtype=np.int32)
    if world == 1:
        if timer is not None:
            h.timer_start()
        out[:] = h.gccb_simulate(Um, s, n_samples, eta=eta, seed=seed)
        if timer is not None:
            timer.append(h.timer_stop())
        return out
    store = dist.distributed_c10d._get_default_store()
    key = f"bossperm/dynamic/{_dynamic_calls}"
    ms, mine = 0.0, 0
    while True:
        hi = int(store.add(key, int(batch)))            # atomic: the batch [hi - batch, hi) is this rank's
        lo = hi - int(batch)
        if lo >= n_samples:
            break
        hi = min(hi, int(n_samples))
        if timer is not None:
            h.timer_start()
        out[lo:hi] = h.gccb_simulate(Um, s, hi - lo, eta=eta, seed=seed, first_sample=lo)
        if timer is not None:
            ms += h.timer_stop()
        mine += hi - lo
    if timer is not None:
        timer.append(ms)
        timer.append(mine)
    on_gpu = dist.get_backend(group) != "gloo"          # (gloo: the CPU test of this scheduling logic)
    t = torch.from_numpy(out).to(torch.device("cuda", dev)) if on_gpu else torch.from_numpy(out)
    dist.all_reduce(t, group=group)                     # every row was written by exactly one rank, the others hold zeros
    return t.cpu().numpy()

"""Multi-GPU plumbing: one process per GPU, evidence cases sharded by contiguous ranges.

Cases are independent (belief_propagation.hpp keeps all state per call, :162-172), so there is NO
data-path collective.  The only exchanges are the two the north star names: the global convergence
summary (sum of sweeps, all-converged flag) and, on demand, the gather of posterior marginals.
Works with any ``torch.distributed`` backend (``nccl`` on the GPU box, ``gloo`` in the CPU tests).

Since round 2 the PRODUCT's exchanges live inside libbnbp (``bnbp_comm_*``, ``bnbp_create_multi``, the chunked gather
overlapped with the kernels: ``engine.BeliefPropagation.comm_init`` / ``run_device(gather=True)`` / ``comm_summary``),
where a C++ host can reach them.  What stays here is the sharding arithmetic (the same ranges the library cuts,
``shard_range``) and torch-level equivalents of the two exchanges, which the world-size-2 gloo tests use to check the
host logic on a machine without GPUs (``tests/test_dist_gloo.py``)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

from .flat import EvidenceBatch


def shard_range(n_cases: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank ``rank``: ranges differ by at most one case."""
    base, rem = divmod(n_cases, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_evidence(ev: EvidenceBatch, rank: int, world: int) -> EvidenceBatch:
    lo, hi = shard_range(ev.n_cases, rank, world)
    return ev.slice(lo, hi)


def reduce_summary(sweeps: torch.Tensor, converged: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[0] = sum over ALL ranks of per-case sweeps, out[1] = number of non-converged cases.
    One 16-byte all-reduce; asynchronous w.r.t. the host on CUDA tensors."""
    out[0] = sweeps.sum(dtype=torch.int64)
    out[1] = (converged == 0).sum(dtype=torch.int64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def gather_marginals(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather case-major marginals [n_local, V] into [n_total, V] in global case order.
    Shards may differ by one case (shard_range), so they are padded to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    V = local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == mx for lo, hi in sizes):
        out = torch.empty((world * mx, V), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((mx, V), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * mx, V), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)

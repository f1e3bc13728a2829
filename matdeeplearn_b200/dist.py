"""Data parallelism for the engine: whole graphs sharded across ranks, ONE
all-reduce of a single contiguous fp32 gradient buffer per step.

Replaces the reference's DistributedDataParallel wrap
(matdeeplearn/training/training.py:262-266: bucketed mean-allreduce with
find_unused_parameters=True, BN buffers broadcast every forward) and its
DistributedSampler sharding (training.py:291-294).  Semantics kept: gradients
are averaged over ranks; BatchNorm statistics stay per-rank (plain DDP, no
SyncBN); per-rank batch size is fixed (weak scaling), lr is scaled by the world
size by the caller (training.py:389).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatParameters:
    """Re-homes every parameter of `module` (and its gradient) as a view into one
    contiguous fp32 buffer each, so the data-parallel exchange is a single
    collective and the optimizer a single fused kernel."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, "module has no trainable parameters"
        dev, dt = params[0].device, params[0].dtype
        assert all(p.dtype == dt and p.device == dev for p in params)
        total = sum(p.numel() for p in params)
        self.param = torch.empty(total, dtype=dt, device=dev)
        self.grad = torch.zeros(total, dtype=dt, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            self.param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.param[off:off + n].view_as(p)
            p.grad = self.grad[off:off + n].view_as(p)
            off += n
        self.params = params
        self.numel = total
        # a leaf the optimizer can own: same storage as every view above
        self.leaf = torch.nn.Parameter(self.param, requires_grad=True)
        self.leaf.grad = self.grad

    def zero_grad(self):
        self.grad.zero_()

    def bind_grads(self):
        """Re-attach gradient views (needed if someone set .grad to None)."""
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.grad[off:off + n].view_as(p)
            off += n


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_mean_(flat_grad: torch.Tensor):
    """In-place mean over ranks of the flat gradient buffer (DDP semantics)."""
    if not is_distributed():
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    flat_grad.mul_(1.0 / dist.get_world_size())
    return flat_grad


def broadcast_(flat: torch.Tensor, src=0):
    """Rank-0 parameters to everyone at start-up (what DDP's constructor does)."""
    if is_distributed():
        dist.broadcast(flat, src=src)
    return flat


def shard_indices(num_items: int, rank: int, world: int, sizes=None):
    """Whole-graph sharding.  Without `sizes`: strided like DistributedSampler
    (rank, rank+world, ...).  With per-graph `sizes` (edge counts): greedy
    longest-first balancing so MOF-shaped batches load every GPU evenly."""
    if sizes is None:
        return list(range(rank, num_items, world))
    order = sorted(range(num_items), key=lambda i: -int(sizes[i]))
    loads = [0] * world
    buckets = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: loads[j])
        buckets[r].append(i)
        loads[r] += int(sizes[i])
    return sorted(buckets[rank])

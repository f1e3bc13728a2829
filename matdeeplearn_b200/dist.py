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
        self.direct = False
        # a leaf the optimizer can own: same storage as every view above
        self.leaf = torch.nn.Parameter(self.param, requires_grad=True)
        self.leaf.grad = self.grad

    def zero_grad(self):
        self.grad.zero_()

    # ---- cheaper per-step protocol used by engine.TrainStep: instead of zeroing the flat gradient
    # and letting autograd ADD into ~20 views (one small kernel each), gradients are detached
    # (autograd then just hands over its result tensors) and packed with ONE concatenation.
    def enable_direct(self):
        """Direct delivery: every parameter advertises its slice of the flat gradient buffer
        (`_mdl_grad_dest`); the engine's autograd Functions (functional.CGConvFn / LinearFn /
        MaskedBatchNormFn) then write weight and bias gradients straight into it and return None,
        so nothing is concatenated afterwards.  Parameters whose gradient still arrives through
        autograd are copied in by pack_grads()."""
        off = 0
        for p in self.params:
            n = p.numel()
            p._mdl_grad_dest = self.grad[off:off + n].view_as(p)
            p._mdl_written = False
            off += n
        self.direct = True

    def release_grads(self):
        for p in self.params:
            p.grad = None
            if self.direct:
                p._mdl_written = False

    def pack_grads(self):
        if not self.direct:
            flat = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
            torch.cat(flat, out=self.grad)
            return self.grad
        dst, src, extra = [], [], []
        for p in self.params:
            if p._mdl_written:
                if p.grad is not None:          # a second use of the parameter went through autograd: add
                    extra.append((p._mdl_grad_dest, p.grad))
            elif p.grad is not None:
                dst.append(p._mdl_grad_dest)
                src.append(p.grad)
            else:
                p._mdl_grad_dest.zero_()
        if dst:
            torch._foreach_copy_(dst, src)
        for d, g in extra:
            d.add_(g)
        return self.grad

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


class FlatAdamW:
    """torch.optim.AdamW semantics over FlatParameters through mdl_adamw_step (one elementwise
    kernel over the whole model).  Hyper-parameters live in a device tensor so a scheduler
    (the reference uses ReduceLROnPlateau, training.py:433-436) can change lr between replays
    of a captured step."""

    def __init__(self, flat: FlatParameters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        from . import _lib
        self._lib = _lib
        self.flat = flat
        dev = flat.param.device
        self.exp_avg = torch.zeros_like(flat.param)
        self.exp_avg_sq = torch.zeros_like(flat.param)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)

    def set_lr(self, lr):
        self.hyper[0] = float(lr)

    def step(self, grad_scale=1.0):
        L, P = self._lib.load(), self._lib.ptr
        rc = L.mdl_adamw_step(P(self.flat.param), P(self.flat.grad), P(self.exp_avg), P(self.exp_avg_sq),
                              P(self.hyper), P(self.step_count), float(grad_scale), self.flat.numel,
                              self._lib.stream())
        self._lib.check(rc, "mdl_adamw_step")

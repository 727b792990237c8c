"""Batch-sharded data parallelism for the adversarial step.

The reference is single-process / single-device.  The hot path shards naturally by batch: rank r
of R takes samples [r*B/R, (r+1)*B/R) of every tensor.  Entropy map and Chamfer need no data-path
collective; the only exchange is the all-reduce (mean) of D4's parameter gradients before its SGD
step, done here as ONE flat bucket so the cost is one latency-bound NCCL call over NVLink/NVSwitch.
torch.distributed is plumbing (process group, NCCL communicator); works with gloo on CPU for tests.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of `batch` samples; the first (batch % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    r = shard_range(t.shape[0], rank, world)
    return t[r.start:r.stop]


def check_per_rank_batch(batch: int, world: int) -> None:
    """D4 cannot run a per-rank batch of 1 (the reference's B == 1 branch is broken)."""
    if batch // world < 2:
        raise ValueError(f"global batch {batch} over {world} ranks leaves < 2 samples per rank; "
                         "PointNetCls needs >= 2 clouds per call")


class GradBucket:
    """One flat fp32 buffer aliasing the gradients of a parameter list; all-reduced in one call."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("empty parameter list")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()

    def attach(self) -> None:
        """Point every .grad at its slice of the bucket (so backward accumulates in place)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self) -> None:
        self.flat.zero_()

    def accumulate(self, grads: Sequence[Optional[torch.Tensor]]) -> None:
        """bucket += grads (as returned by torch.autograd.grad over self.params; None = zero)."""
        vs = [v for v, g in zip(self.views, grads) if g is not None]
        gs = [g.reshape(v.shape) for v, g in zip(self.views, grads) if g is not None]
        if vs:
            torch._foreach_add_(vs, gs)

    def assign_sum(self, grads_a: Sequence[Optional[torch.Tensor]], grads_b: Sequence[Optional[torch.Tensor]]) -> None:
        """bucket = grads_a + grads_b in two multi-tensor launches (no zero fill) when both lists are
        complete; otherwise zero + accumulate."""
        if any(g is None for g in grads_a) or any(g is None for g in grads_b):
            self.zero()
            self.accumulate(grads_a)
            self.accumulate(grads_b)
            return
        torch._foreach_copy_(self.views, [g.reshape(v.shape) for v, g in zip(self.views, grads_a)])
        torch._foreach_add_(self.views, [g.reshape(v.shape) for v, g in zip(self.views, grads_b)])

    def gather_grads(self) -> None:
        """Copy .grad tensors that are not already views of the bucket into it."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def allreduce_sum(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """Sum over ranks (for a bucket that was filled pre-divided by the world size)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def allreduce_mean(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        if world == 1:
            return None
        self.flat.div_(world)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def allreduce_scalars(values: Sequence[torch.Tensor], group=None) -> torch.Tensor:
    """Mean over ranks of a handful of logged scalars (losses) — one tiny collective."""
    t = torch.stack([v.detach().reshape(()) for v in values])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t / dist.get_world_size(group)
    return t


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group if
    WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world

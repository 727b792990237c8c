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


class _DeviceSpan:
    """A span of device memory owned by libpcuda, exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, numel: int, owner):
        self.owner = owner        # keeps the communicator (and with it the memory) alive as long as a view exists
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PcudaComm:
    """libpcuda's communicator (include/pcuda.h `pcuda_comm_*`): an NCCL communicator bound at run time plus, for up
    to 8 ranks of one node with peer access, a symmetric peer-memory region for the NVLink all-reduce kernel.

    Collective constructor: every rank of `group` calls it.  torch.distributed is used once, to hand rank 0's
    128-byte NCCL id to the other ranks."""

    def __init__(self, device: torch.device, p2p_floats: int = 0, group: Optional[dist.ProcessGroup] = None):
        import ctypes
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PcudaComm needs an initialised torch.distributed process group (plumbing for the id exchange)")
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        buf = ctypes.create_string_buffer(128)
        with torch.cuda.device(self.device):
            if self.rank == 0:
                _lib.check(L.pcuda_comm_unique_id(buf, 128), "pcuda_comm_unique_id")
            box = [bytes(buf.raw)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            idbuf = ctypes.create_string_buffer(box[0], 128)
            handle = ctypes.c_void_p()
            _lib.check(L.pcuda_comm_init(idbuf, self.rank, self.world, int(p2p_floats), ctypes.byref(handle)), "pcuda_comm_init")
        self._h = handle
        r, w, p2p, ver = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(L.pcuda_comm_info(self._h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(p2p), ctypes.byref(ver)), "pcuda_comm_info")
        self.p2p, self.nccl_version = bool(p2p.value), int(ver.value)
        self.buf_in = self.buf_out = None
        if self.p2p:
            pin, pout, cap = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
            _lib.check(L.pcuda_comm_p2p_buffers(self._h, ctypes.byref(pin), ctypes.byref(pout), ctypes.byref(cap)), "pcuda_comm_p2p_buffers")
            self.capacity = int(cap.value)
            self.buf_in = torch.as_tensor(_DeviceSpan(pin.value, self.capacity, self), device=self.device)
            self.buf_out = torch.as_tensor(_DeviceSpan(pout.value, self.capacity, self), device=self.device)

    def allreduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks (ncclAllReduce on the current stream; CUDA-graph capturable)."""
        from . import _lib
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
            raise _lib.PcudaError("PcudaComm.allreduce_ needs a contiguous float32 tensor on the communicator's device")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pcuda_comm_allreduce(self._h, t.data_ptr(), t.numel(), torch.cuda.current_stream(self.device).cuda_stream),
                       "pcuda_comm_allreduce")
        return t

    def allreduce_f64_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum of a float64 tensor over the ranks (the exchange of BatchNorm sums in the cross-rank mode): up to
        2048 values go through the one-shot peer-memory mailbox kernel, more through ncclAllReduce."""
        from . import _lib
        if t.dtype != torch.float64 or not t.is_contiguous() or t.device != self.device:
            raise _lib.PcudaError("PcudaComm.allreduce_f64_ needs a contiguous float64 tensor on the communicator's device")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pcuda_comm_allreduce_f64(self._h, t.data_ptr(), t.numel(), torch.cuda.current_stream(self.device).cuda_stream),
                       "pcuda_comm_allreduce_f64")
        return t

    def allgather(self, send: torch.Tensor, recv: torch.Tensor) -> torch.Tensor:
        """recv[r * n:(r + 1) * n] = rank r's send (n = send.numel(); ncclAllGather on the current stream)."""
        from . import _lib
        if send.dtype != torch.float32 or recv.dtype != torch.float32 or not send.is_contiguous() or not recv.is_contiguous() \
                or recv.numel() != self.world * send.numel():
            raise _lib.PcudaError("PcudaComm.allgather needs contiguous float32 tensors, recv = world x send")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pcuda_comm_allgather(self._h, send.data_ptr(), recv.data_ptr(), send.numel(),
                                                       torch.cuda.current_stream(self.device).cuda_stream), "pcuda_comm_allgather")
        return recv

    def allreduce_p2p(self, count: int) -> torch.Tensor:
        """Sum of every rank's `buf_in[:count]` into every rank's `buf_out[:count]` (one NVLink peer-memory kernel)."""
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pcuda_comm_allreduce_p2p(self._h, int(count), torch.cuda.current_stream(self.device).cuda_stream),
                       "pcuda_comm_allreduce_p2p")
        return self.buf_out[:count]

    def check_status(self) -> None:
        from . import _lib
        _lib.check(_lib.lib().pcuda_comm_status(self._h), "pcuda_comm_status")

    def destroy(self) -> None:
        from . import _lib
        if self._h is not None and self._h.value:
            self.buf_in = self.buf_out = None
            _lib.lib().pcuda_comm_destroy(self._h)
            self._h = None


class GradBucket:
    """One flat fp32 buffer aliasing the gradients of a parameter list; all-reduced in one call."""

    def __init__(self, params: Iterable[torch.nn.Parameter], flat: Optional[torch.Tensor] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("empty parameter list")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        if flat is not None:        # e.g. the head of a communicator's peer-memory input buffer
            if flat.numel() != self.numel or flat.dtype != dt or flat.device != dev or not flat.is_contiguous():
                raise ValueError("flat buffer does not match the parameter list")
            self.flat = flat
            self.flat.zero_()
        else:
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


def bind_to_gpu_numa(local_rank: int) -> Optional[str]:
    """Pin this process (and with it the pages of the pinned host buffers it allocates afterwards) to the CPU cores
    NVML reports as local to GPU `local_rank`.  With one process per GPU feeding inputs from pinned host memory, all
    ranks otherwise share whatever NUMA node the scheduler happened to start them on.  Returns a description, or None
    when NVML / affinity control is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cores local to GPU {local_rank}"
    except Exception:
        return None


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

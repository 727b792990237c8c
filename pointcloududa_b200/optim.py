"""D4's gradient bucket fill and SGD step as one libpcuda launch each (csrc/optim.cu).

Mirrors `optim_dis4 = torch.optim.SGD(model_dis4.parameters(), lr, momentum=.99, weight_decay=.0005)` and its
`.step()` (train_mscmrseg.py:329-330,:450-455): same arithmetic and order as torch's SGD (dampening 0, nesterov
off).  The momentum buffers live in one flat tensor (same layout as dist.GradBucket.flat), the learning rate in a
1-element device tensor so a schedule (`adjust_learning_rate`, utils/utils.py) never invalidates a captured graph.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import ParamSlot, PcudaError, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class FusedSGD:
    """SGD with momentum and weight decay over a fixed parameter list, gradients read from a flat bucket."""

    def __init__(self, params: Sequence[torch.nn.Parameter], lr: float, momentum: float = 0.0, weight_decay: float = 0.0):
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise PcudaError("FusedSGD needs CUDA parameters (libpcuda has no CPU path)")
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise PcudaError("FusedSGD needs contiguous float32 parameters on one device")
        self.device = dev
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.numel = sum(p.numel() for p in self.params)
        self.momentum_buffer = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._slots = (ParamSlot * len(self.params))()
        off = 0
        for s, p in zip(self._slots, self.params):
            s.grad_a, s.grad_b, s.param, s.offset, s.numel = None, None, p.data_ptr(), off, p.numel()
            off += p.numel()

    def set_lr(self, lr: float) -> None:
        self.lr.fill_(float(lr))

    def pack(self, grads_a: Sequence[torch.Tensor], grads_b: Optional[Sequence[torch.Tensor]], flat: torch.Tensor,
             scale: float = 1.0) -> None:
        """flat[slice_i] = scale * (grads_a[i] + grads_b[i]) for every parameter, one launch."""
        if flat.numel() != self.numel or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise PcudaError("pack: flat bucket does not match the parameter list")
        keep = self._bind_grads(grads_a, grads_b)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            check(L.pcuda_grad_sum_pack(self._slots, len(self._slots), float(scale), flat.data_ptr(), st), "pcuda_grad_sum_pack")

    def _bind_grads(self, grads_a, grads_b) -> list:
        keep = []
        for i, s in enumerate(self._slots):
            ga = grads_a[i]
            gb = grads_b[i] if grads_b is not None else None
            if ga is None or (grads_b is not None and gb is None):
                raise PcudaError("pack: missing gradient (use GradBucket.accumulate for partial lists)")
            ga = ga.contiguous()
            gb = gb.contiguous() if gb is not None else None
            if ga.numel() != s.numel or (gb is not None and gb.numel() != s.numel):
                raise PcudaError("pack: gradient shape mismatch")
            keep += [ga, gb]
            s.grad_a, s.grad_b = ga.data_ptr(), _ptr(gb)
        return keep

    def sum_step(self, grads_a: Sequence[torch.Tensor], grads_b: Optional[Sequence[torch.Tensor]], flat: torch.Tensor,
                 scale: float = 1.0) -> None:
        """pack() followed by step(flat) in ONE launch: for a single process, where nothing is exchanged in between.
        `flat` still receives the summed gradient."""
        if flat.numel() != self.numel or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise PcudaError("sum_step: flat bucket does not match the parameter list")
        keep = self._bind_grads(grads_a, grads_b)
        for s, p in zip(self._slots, self.params):
            s.param = p.data_ptr()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            check(L.pcuda_sgd_momentum_sum_step(self._slots, len(self._slots), float(scale), flat.data_ptr(),
                                                self.momentum_buffer.data_ptr(), self.lr.data_ptr(), self.momentum,
                                                self.weight_decay, st), "pcuda_sgd_momentum_sum_step")
        del keep

    def step(self, flat_grad: torch.Tensor) -> None:
        """One SGD step of every parameter from the flat gradient bucket, one launch."""
        if flat_grad.numel() != self.numel or flat_grad.dtype != torch.float32 or not flat_grad.is_contiguous():
            raise PcudaError("step: flat gradient does not match the parameter list")
        for s, p in zip(self._slots, self.params):
            s.param = p.data_ptr()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            check(L.pcuda_sgd_momentum_step(self._slots, len(self._slots), flat_grad.data_ptr(), self.momentum_buffer.data_ptr(),
                                            self.lr.data_ptr(), self.momentum, self.weight_decay, st), "pcuda_sgd_momentum_step")

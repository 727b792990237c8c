"""torch.autograd wrappers over the libpcuda C-ABI.

PyTorch is plumbing here: it owns device memory, streams and the autograd graph; every number is
produced by a kernel in csrc/.  All functions require CUDA fp32 tensors and raise otherwise —
there is no CPU path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import PcudaError, check

ACT = {"sigmoid": 0, "softmax": 1}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise PcudaError(f"{name} is on {t.device}: pointcloududa_b200 runs on CUDA (sm_100a) only; "
                         "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise PcudaError(f"{name} has dtype {t.dtype}; the hot path computes in fp32")


# Small zero-initialised workspaces the kernels leave zeroed (tickets / partial sums), cached per
# (device, stream) so concurrent streams never share one.
_ws_cache: dict = {}
_ws_retired: list = []      # superseded workspaces are never freed: a captured CUDA graph or in-flight work may hold the pointer


def _zero_ws(nbytes: int, device: torch.device, tag: str) -> torch.Tensor:
    key = (tag, device.index, _stream())
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _ws_retired.append(t)
        t = torch.zeros(max(nbytes, 4096), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def _check_ws(rc: int, what: str, ws: Optional[torch.Tensor]) -> None:
    """check(), and on failure re-zero the ticket workspace: a rejected or failed launch may leave partial sums / a
    ticket behind, after which no later call on this stream would elect its last CTA."""
    try:
        check(rc, what)
    except PcudaError:
        if ws is not None:
            ws.zero_()
        raise


# ------------------------------------------------------------------------------------------------
class _EntropyMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, activation: str, normalize: bool, smooth: float, want_p: bool, want_mean: bool):
        _require_cuda_f32(z, "logits")
        if z.dim() < 2:
            raise PcudaError("logits must be [B, C, ...]")
        zc = z.contiguous()
        B, C = zc.shape[0], zc.shape[1]
        HW = zc.numel() // max(B * C, 1) if B * C > 0 else 0
        inv_norm = 1.0 / math.log(C) if normalize else 1.0
        if normalize and C < 2:
            raise PcudaError("normalize=True needs C >= 2 (division by log C)")
        m = torch.empty_like(zc)
        p = torch.empty_like(zc) if want_p else None
        mean = torch.empty((), dtype=torch.float32, device=zc.device) if want_mean else None
        with torch.cuda.device(zc.device):
            ws = _zero_ws(16, zc.device, "entropy") if want_mean else None
            _check_ws(_lib.lib().pcuda_entropy_fwd(_ptr(zc), _ptr(m), _ptr(p), _ptr(mean), _ptr(ws), B, C, HW,
                                                  ACT[activation], inv_norm, smooth, _stream()),
                      "pcuda_entropy_fwd", ws)
        ctx.save_for_backward(zc)
        ctx.cfg = (activation, inv_norm, smooth, B, C, HW)
        return m, p, mean

    @staticmethod
    def backward(ctx, gm, gp, gmean):
        (zc,) = ctx.saved_tensors
        activation, inv_norm, smooth, B, C, HW = ctx.cfg
        gm = None if gm is None else gm.contiguous()
        gp = None if gp is None else gp.contiguous()
        gmean = None if gmean is None else gmean.contiguous().to(torch.float32)
        gz = torch.empty_like(zc)
        with torch.cuda.device(zc.device):
            check(_lib.lib().pcuda_entropy_bwd(_ptr(zc), _ptr(gm), _ptr(gp), _ptr(gmean), _ptr(gz), B, C, HW,
                                              ACT[activation], inv_norm, smooth, _stream()),
                  "pcuda_entropy_bwd")
        return gz, None, None, None, None, None


def entropy_map(logits: torch.Tensor, activation: str = "sigmoid", normalize: bool = False,
                smooth: float = 1e-7, return_prob: bool = False, return_mean: bool = False):
    """Fused self-information map  m = -p*log(p+smooth) [/ log C].

    Replaces the inline expressions at train_mscmrseg.py:222,265 (activation='sigmoid',
    normalize=False) and train_mmwhs.py:213-217,224-225,240-243 (activation per -softmax,
    normalize=True, return_prob=True for D1's input, return_mean=True for the logged
    mean(sum(map, dim=1))).  Returns m, or a tuple (m[, p][, mean]) in that order.
    """
    if activation not in ACT:
        raise ValueError(f"activation must be 'sigmoid' or 'softmax', got {activation!r}")
    m, p, mean = _EntropyMap.apply(logits, activation, bool(normalize), float(smooth),
                                   bool(return_prob), bool(return_mean))
    out = [m]
    if return_prob:
        out.append(p)
    if return_mean:
        out.append(mean)
    return out[0] if len(out) == 1 else tuple(out)


# ------------------------------------------------------------------------------------------------
class _ChamferNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        _require_cuda_f32(x, "x")
        _require_cuda_f32(y, "y")
        if x.dim() != 3 or y.dim() != 3 or x.shape[2] != 3 or y.shape[2] != 3:
            raise PcudaError(f"expected x [B,N,3] and y [B,M,3], got {tuple(x.shape)} and {tuple(y.shape)}")
        if x.shape[0] != y.shape[0]:
            raise PcudaError("x and y must have the same batch size")
        if x.device != y.device:
            raise PcudaError("x and y must be on the same device")
        xc, yc = x.contiguous(), y.contiguous()
        B, N, M = xc.shape[0], xc.shape[1], yc.shape[1]
        dev = xc.device
        d1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        i1 = torch.empty((B, N), dtype=torch.int64, device=dev)
        d2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        i2 = torch.empty((B, M), dtype=torch.int64, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L = _lib.lib()
        with torch.cuda.device(dev):
            ws = _zero_ws(L.pcuda_chamfer_ws_bytes(B, N, M), dev, "chamfer")
            _check_ws(L.pcuda_chamfer_fwd(_ptr(xc), _ptr(yc), B, N, M, _ptr(d1), _ptr(i1), _ptr(d2), _ptr(i2),
                                          _ptr(loss), _ptr(ws), _stream()), "pcuda_chamfer_fwd", ws)
        ctx.save_for_backward(xc, yc, d1, i1, d2, i2)
        ctx.mark_non_differentiable(d1, i1, d2, i2)
        return loss, d1, i1, d2, i2

    @staticmethod
    def backward(ctx, gloss, *_):
        xc, yc, d1, i1, d2, i2 = ctx.saved_tensors
        B, N, M = xc.shape[0], xc.shape[1], yc.shape[1]
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = torch.empty_like(xc) if need_x else None
        gy = torch.empty_like(yc) if need_y else None
        if not (need_x or need_y):
            return None, None
        gl = gloss.contiguous().to(torch.float32)
        L = _lib.lib()
        with torch.cuda.device(xc.device):
            ws = torch.empty(max(L.pcuda_chamfer_bwd_ws_bytes(B, N, M), 16), dtype=torch.uint8, device=xc.device)
            check(L.pcuda_chamfer_bwd(_ptr(xc), _ptr(yc), _ptr(d1), _ptr(i1), _ptr(d2), _ptr(i2), _ptr(gl),
                                      B, N, M, _ptr(gx), _ptr(gy), _ptr(ws), _stream()), "pcuda_chamfer_bwd")
        return gx, gy


def chamfer_nn(x: torch.Tensor, y: torch.Tensor):
    """Returns (loss, d1, i1, d2, i2): the scalar of batch_NN_loss plus the per-point nearest
    distances / first-arg-min indices the reference computes and discards (loss.py:68-72)."""
    return _ChamferNN.apply(x, y)


class _BceLogits(torch.autograd.Function):
    """weight * F.binary_cross_entropy_with_logits(logit, full_like(logit, target)) (reduction 'mean'), its gradient
    and the discriminator accuracy in one launch (pcuda_bce_logits)."""

    @staticmethod
    def forward(ctx, logit, target: float, weight: float):
        _require_cuda_f32(logit, "logit")
        x = logit.contiguous()
        dev = x.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        acc = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty_like(x)
        with torch.cuda.device(dev):
            check(_lib.lib().pcuda_bce_logits(_ptr(x), x.numel(), float(target), float(weight), _ptr(loss), _ptr(grad),
                                              _ptr(acc), _stream()), "pcuda_bce_logits")
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, gloss, _gacc):
        (grad,) = ctx.saved_tensors
        return grad * gloss, None, None


def bce_with_logits(logit: torch.Tensor, target: float, weight: float = 1.0):
    """Returns (loss, accuracy): loss = weight * mean BCE-with-logits against the constant label `target`
    (train_mscmrseg.py:233,286,316), accuracy = fraction of samples whose sigmoid(logit) >= 0.5 agrees with the label
    (train_mscmrseg.py:290-296,:320-322)."""
    return _BceLogits.apply(logit, float(target), float(weight))


def bce_with_logits_grad(logit: torch.Tensor, target: float, weight: float = 1.0):
    """(loss, accuracy, d loss / d logit) of bce_with_logits in the same one launch, outside autograd: the caller
    continues with torch.autograd.grad(logit, inputs, grad_outputs=grad) — no unit root gradient to fill, no
    gradient-times-one launch on the way back (the adversarial step's three D4 passes end this way)."""
    _require_cuda_f32(logit, "logit")
    x = logit.detach().contiguous()
    dev = x.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    acc = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(x)
    with torch.cuda.device(dev):
        check(_lib.lib().pcuda_bce_logits(_ptr(x), x.numel(), float(target), float(weight), _ptr(loss), _ptr(grad),
                                          _ptr(acc), _stream()), "pcuda_bce_logits")
    return loss, acc, grad


def launch_count() -> int:
    """Kernels launched by libpcuda in this process so far."""
    return int(_lib.lib().pcuda_launch_count())


def sm_count() -> int:
    """SM count of the current device (148 on B200)."""
    return int(_lib.lib().pcuda_sm_count())


def tune(key: int, value: int) -> None:
    check(_lib.lib().pcuda_tune(int(key), int(value)), "pcuda_tune")

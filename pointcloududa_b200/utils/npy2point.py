"""Drop-in for the farthest-point sampling of the reference's `src/utils/npy2point.py` (graipher, :11-18), on the GPU.

`graipher(pts, K, dim=2)` keeps the reference's signature and semantics: the first point is drawn with
`np.random.randint(len(pts))` (the same draw from numpy's global generator as the reference, so a seeded data pipeline
selects the same points), the other K - 1 greedily by largest squared distance to the chosen set, float64 arithmetic,
`np.argmax` tie-breaking.  The selection runs in one libpcuda launch (csrc/fps.cu); numpy in -> numpy out (the data
generators call it from the loader thread), CUDA tensors in -> CUDA tensors out.  `graipher_batch` samples many clouds
in one launch.  Marching cubes (PyMCubes, `npy2point_datagenerator` :101-125) stays where it is.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .._lib import PcudaError, check


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise PcudaError("farthest-point sampling runs on CUDA (sm_100a) only; there is no CPU fallback")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def graipher_batch(pts: torch.Tensor, K: int, counts: Optional[torch.Tensor] = None, starts: Optional[torch.Tensor] = None):
    """pts [B, V, dim] float64 CUDA, counts [B] int32 (valid points per cloud), starts [B] int32 -> (points [B, K, dim]
    float64, indices [B, K] int32)."""
    if not pts.is_cuda or pts.dtype != torch.float64 or pts.dim() != 3:
        raise PcudaError("graipher_batch expects a float64 CUDA tensor [B, V, dim]")
    B, V, dim = pts.shape
    p = pts.contiguous()
    out = torch.empty((B, K, dim), dtype=torch.float64, device=p.device)
    idx = torch.empty((B, K), dtype=torch.int32, device=p.device)
    c = None if counts is None else counts.to(device=p.device, dtype=torch.int32).contiguous()
    s = None if starts is None else starts.to(device=p.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(p.device):
        check(_lib.lib().pcuda_fps(p.data_ptr(), None if c is None else c.data_ptr(), None if s is None else s.data_ptr(), B, V, int(K),
                                   dim, out.data_ptr(), idx.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream), "pcuda_fps")
    return out, idx


def graipher(pts, K, dim=2, start: Optional[int] = None, device=None):
    """utils/npy2point.py:11-18.  pts: [V, dim] numpy array (returns a numpy array [K, dim] like the reference) or CUDA
    tensor (returns a CUDA tensor).  start: index of the first point; None draws np.random.randint(len(pts)) like the
    reference."""
    if len(pts) == 0:
        raise ValueError("graipher: empty point set (the reference raises in np.random.randint)")
    if start is None:
        start = int(np.random.randint(len(pts)))
    as_numpy = not isinstance(pts, torch.Tensor)
    t = torch.as_tensor(np.asarray(pts, dtype=np.float64)) if as_numpy else pts.to(torch.float64)
    if t.dim() != 2 or t.shape[1] < dim:
        raise ValueError(f"graipher: expected [V, >= {dim}] points, got {tuple(t.shape)}")
    t = t[:, :dim].to(_device(device) if as_numpy else t.device)
    out, _ = graipher_batch(t[None], int(K), starts=torch.tensor([start], dtype=torch.int32))
    return out[0].cpu().numpy() if as_numpy else out[0]

"""Seeded synthetic inputs of the reference's shapes.

Generated on the host with a torch CPU generator so the CPU oracle, the golden fixtures and the
CUDA path all see identical bytes.  Datasets are not shipped with the reference, so these mimic
its value distributions:
  logits      N(0, 3^2) fp32 NCHW
  GT clouds   (randint(0,3), randint(0,256), randint(0,256)) / 255 — marching-cubes vertices cast
              to int then /255 (utils/npy2point.py:70-81,124; data_generator_mscmrseg.py:317), with
              one all-zero cloud per batch (empty masks, npy2point.py:72,115)
  predicted   y[perm] + 0.02*N(0,1)   (a segmenter that is nearly right), or U[0,1)^3
"""
from __future__ import annotations

import torch

# BASELINE.json configs (SURVEY.md §8a shorthand)
CONFIGS = {
    "cfg1": dict(B=8, C=4, H=224, W=224, N=300, activation="sigmoid", normalize=False),
    "cfg2": dict(B=8, C=4, H=224, W=224, N=300, activation="sigmoid", normalize=False),
    "cfg3": dict(B=32, C=5, H=256, W=256, N=1024, activation="softmax", normalize=True),
    "cfg5": dict(B=256, C=5, H=256, W=256, N=4096, activation="softmax", normalize=True),
}


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def synth_logits(B: int, C: int, H: int, W: int, seed: int = 1234, std: float = 3.0) -> torch.Tensor:
    return torch.randn(B, C, H, W, generator=_gen(seed), dtype=torch.float32) * std


def synth_clouds(B: int, N: int, seed: int = 1234, mode: str = "near", zero_cloud: bool = True):
    """Returns (x_pred [B,N,3], y_gt [B,N,3]) fp32."""
    g = _gen(seed)
    y = torch.stack([
        torch.randint(0, 3, (B, N), generator=g),
        torch.randint(0, 256, (B, N), generator=g),
        torch.randint(0, 256, (B, N), generator=g),
    ], dim=-1).to(torch.float32) / 255.0
    if zero_cloud and B > 0:
        y[B // 2] = 0.0
    if mode == "near":
        perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
        x = torch.gather(y, 1, perm[:, :, None].expand(B, N, 3)) + 0.02 * torch.randn(B, N, 3, generator=g)
    elif mode == "uniform":
        x = torch.rand(B, N, 3, generator=g)
    else:
        raise ValueError(mode)
    return x.contiguous(), y.contiguous()

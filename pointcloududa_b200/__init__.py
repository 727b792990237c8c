"""pointcloududa_b200 — B200-native (sm_100a) kernels for PointCloudUDA's adversarial-adaptation
hot path, behind the reference's own call signatures.

    from pointcloududa_b200.utils.loss import batch_NN_loss, entropy_map
    from pointcloududa_b200.networks.PointNetCls import PointNetCls

Everything computes in hand-written CUDA (csrc/, C-ABI in include/pcuda.h); there is no CPU or
eager-PyTorch fallback — a missing libpcuda.so or a CPU tensor raises.
"""
__version__ = "0.1.0"

__all__ = ["batch_NN_loss", "entropy_map", "chamfer_nn", "PointNetCls", "shared_mlp", "PcudaError"]


def __getattr__(name):
    # lazy: importing the package must not import torch or load the library (build tooling uses it)
    if name in ("batch_NN_loss",):
        from .utils.loss import batch_NN_loss
        return batch_NN_loss
    if name in ("entropy_map", "chamfer_nn"):
        from . import ops
        return getattr(ops, name)
    if name in ("PointNetCls", "shared_mlp"):
        from .networks import PointNetCls as _m
        return getattr(_m, name)
    if name == "PcudaError":
        from ._lib import PcudaError
        return PcudaError
    raise AttributeError(name)

"""ctypes binding of libpcuda.so (the C-ABI declared in include/pcuda.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception
is raised — nothing in this package computes the hot path any other way.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libpcuda.so"
_lib = None


class PcudaError(RuntimeError):
    pass


class MlpLayer(ctypes.Structure):
    """pcuda_mlp_layer_t (include/pcuda.h)."""
    _fields_ = [
        ("cin", c_int32), ("cout", c_int32), ("relu", c_int32), ("reserved", c_int32),
        ("weight", c_void_p), ("bias", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
        ("running_mean", c_void_p), ("running_var", c_void_p),
        ("save_mean", c_void_p), ("save_invstd", c_void_p), ("y", c_void_p),
        ("grad_weight", c_void_p), ("grad_bias", c_void_p), ("grad_gamma", c_void_p),
        ("grad_beta", c_void_p),
    ]


class FcLayer(ctypes.Structure):
    """pcuda_fc_layer_t (include/pcuda.h)."""
    _fields_ = [
        ("cin", c_int32), ("cout", c_int32), ("bn", c_int32), ("relu", c_int32),
        ("weight", c_void_p), ("bias", c_void_p), ("mask", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
        ("running_mean", c_void_p), ("running_var", c_void_p),
        ("save_mean", c_void_p), ("save_invstd", c_void_p), ("y", c_void_p), ("a", c_void_p),
        ("grad_weight", c_void_p), ("grad_bias", c_void_p), ("grad_gamma", c_void_p),
        ("grad_beta", c_void_p),
    ]


class BnUpdate(ctypes.Structure):
    """pcuda_bn_update_t (include/pcuda.h)."""
    _fields_ = [
        ("running_mean", c_void_p), ("running_var", c_void_p), ("save_mean", c_void_p), ("save_invstd", c_void_p),
        ("C", c_int32), ("reserved", c_int32), ("count", c_double), ("momentum", c_float), ("eps", c_float),
    ]


class ParamSlot(ctypes.Structure):
    """pcuda_param_slot_t (include/pcuda.h)."""
    _fields_ = [("grad_a", c_void_p), ("grad_b", c_void_p), ("param", c_void_p), ("offset", c_int64), ("numel", c_int64)]


# name -> (restype, argtypes); must list every symbol include/pcuda.h declares
SIGNATURES = {
    "pcuda_version": (c_int, []),
    "pcuda_last_error_string": (c_char_p, []),
    "pcuda_error_name": (c_char_p, [c_int]),
    "pcuda_sm_count": (c_int, []),
    "pcuda_tune": (c_int, [c_int, c_int]),
    "pcuda_launch_count": (ctypes.c_uint64, []),
    "pcuda_entropy_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int64, c_int, c_float, c_float, c_void_p]),
    "pcuda_entropy_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int64, c_int, c_float, c_float, c_void_p]),
    "pcuda_chamfer_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcuda_chamfer_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_chamfer_bwd_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcuda_chamfer_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "pcuda_pointmlp_ws_bytes": (c_size_t, [c_int, c_int, c_int, POINTER(MlpLayer), c_int, c_int]),
    "pcuda_pointmlp_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                   POINTER(MlpLayer), c_int, c_int, c_float, c_float, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_pointmlp_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                   POINTER(MlpLayer), c_int, c_int, c_float, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_pointmlp_bwd_reuse": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                         POINTER(MlpLayer), c_int, c_int, c_float, c_int, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_pointmlp_fwd_xf": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int,
                                      POINTER(MlpLayer), c_int, c_int, c_float, c_float, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_pointmlp_bwd_xf": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int,
                                      POINTER(MlpLayer), c_int, c_int, c_float, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_point_transform_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pcuda_point_transform_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p,
                                          c_void_p]),
    "pcuda_point_transform_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_fcstack_fwd": (c_int, [c_void_p, c_int, c_int, POINTER(FcLayer), c_int, c_float, c_float, c_int,
                                  c_void_p]),
    "pcuda_fcstack_ws_bytes": (c_size_t, [c_int, c_int, POINTER(FcLayer), c_int]),
    "pcuda_fcstack_bwd": (c_int, [c_void_p, c_int, c_int, POINTER(FcLayer), c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "pcuda_bn_running_update": (c_int, [c_int, POINTER(BnUpdate), c_void_p]),
    "pcuda_grad_sum_pack": (c_int, [POINTER(ParamSlot), c_int, c_float, c_void_p, c_void_p]),
    "pcuda_bce_logits": (c_int, [c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pcuda_sgd_momentum_step": (c_int, [POINTER(ParamSlot), c_int, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                        c_void_p]),
    "pcuda_sgd_momentum_sum_step": (c_int, [POINTER(ParamSlot), c_int, c_float, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                            c_void_p]),
    "pcuda_fps": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pcuda_comm_unique_id": (c_int, [c_void_p, c_int]),
    "pcuda_comm_init": (c_int, [c_void_p, c_int, c_int, c_size_t, POINTER(c_void_p)]),
    "pcuda_comm_allreduce": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pcuda_comm_allreduce_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pcuda_comm_allgather": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "pcuda_comm_p2p_buffers": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64)]),
    "pcuda_comm_allreduce_p2p": (c_int, [c_void_p, c_int64, c_void_p]),
    "pcuda_comm_status": (c_int, [c_void_p]),
    "pcuda_comm_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pcuda_comm_destroy": (c_int, [c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise PcudaError(
                f"{_LIB_PATH} not found: build it with `python -m pointcloududa_b200.build` "
                "(nvcc, sm_100a). pointcloududa_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = stale library
            fn.restype = res
            fn.argtypes = args
        if L.pcuda_version() // 100 != 1:
            raise PcudaError(f"libpcuda version {L.pcuda_version()} does not match this package")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        msg = L.pcuda_last_error_string().decode(errors="replace")
        name = L.pcuda_error_name(rc).decode(errors="replace")
        raise PcudaError(f"{what} failed: {name} ({rc}): {msg}")

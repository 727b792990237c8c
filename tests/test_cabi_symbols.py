"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/pcuda.h declares, rejects bad arguments without touching a GPU, and the host-side mirror of
the reference interface keeps the reference's names, signatures and state_dict keys."""
import inspect
import re
from pathlib import Path

import pytest
import torch

from pointcloududa_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pcuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcuda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 14
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), f"libpcuda.so does not export {s}"
        assert s in _lib.SIGNATURES, f"{s} is declared in pcuda.h but has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_names():
    L = _lib.lib()
    assert L.pcuda_version() == 100
    assert L.pcuda_error_name(0) == b"PCUDA_OK"
    assert L.pcuda_error_name(-2) == b"PCUDA_E_SHAPE"


def test_argument_validation_without_gpu():
    L = _lib.lib()
    # bad shapes / unsupported sizes are rejected before any CUDA call
    assert L.pcuda_entropy_fwd(None, None, None, None, None, 1, 0, 16, 0, 1.0, 1e-7, None) == -2
    assert L.pcuda_entropy_fwd(None, None, None, None, None, 1, 17, 16, 1, 1.0, 1e-7, None) == -3
    assert L.pcuda_entropy_fwd(None, None, None, None, None, 1, 4, 16, 7, 1.0, 1e-7, None) == -3
    assert b"activation" in L.pcuda_last_error_string()
    assert L.pcuda_chamfer_fwd(None, None, -1, 4, 4, None, None, None, None, None, None, None) == -2
    assert L.pcuda_chamfer_fwd(None, None, 2, 0, 4, None, None, None, None, None, None, None) == -2
    assert L.pcuda_chamfer_fwd(None, None, 2, 4, 4, None, None, None, None, None, None, None) == -1
    assert L.pcuda_chamfer_ws_bytes(8, 300, 300) >= 16
    assert L.pcuda_chamfer_bwd_ws_bytes(8, 300, 300) == 8 * 3 * 8 * 600
    lay = (_lib.MlpLayer * 1)()
    assert L.pcuda_pointmlp_fwd(None, 0, 0, 0, 2, 8, 1, lay, 1, 1, 0.1, 1e-5, 0, None, None, None, None) == -2
    fc = (_lib.FcLayer * 1)()
    assert L.pcuda_fcstack_fwd(None, 4, 1, fc, 1, 0.1, 1e-5, 0, None) == -2           # cin = cout = 0
    fc[0].cin, fc[0].cout = 6, 4
    assert L.pcuda_fcstack_fwd(None, 4, 1, fc, 1, 0.1, 1e-5, 0, None) == -3           # cin % 4 != 0
    assert L.pcuda_fcstack_fwd(None, 0, 1, fc, 1, 0.1, 1e-5, 0, None) == -2
    fc[0].cin, fc[0].cout = 8, 4
    assert L.pcuda_fcstack_fwd(None, 4, 1, fc, 1, 0.1, 1e-5, 0, None) == -1           # NULL weight
    assert L.pcuda_fcstack_ws_bytes(4, 1, fc, 1) == 2 * 4 * 4 * 4 + 16


def test_cpu_tensors_are_rejected_not_computed():
    """No CPU fallback: the product raises instead of computing on the host."""
    from pointcloududa_b200.utils.loss import batch_NN_loss, entropy_map
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    with pytest.raises(_lib.PcudaError):
        entropy_map(torch.zeros(1, 4, 4, 4))
    with pytest.raises(_lib.PcudaError):
        batch_NN_loss(x=torch.zeros(1, 4, 3), y=torch.zeros(1, 4, 3))
    with pytest.raises(_lib.PcudaError):
        PointNetCls()(torch.zeros(2, 3, 16))
    with pytest.raises(RuntimeError):
        batch_NN_loss(x=torch.zeros(1, 4, 3), y=torch.zeros(1, 5, 3))


def test_reference_interface_mirrored():
    from pointcloududa_b200.utils.loss import batch_NN_loss
    from pointcloududa_b200.networks import PointNetCls as m
    assert list(inspect.signature(batch_NN_loss).parameters) == ["x", "y"]      # utils/loss.py:40
    sig = inspect.signature(m.PointNetCls.__init__)
    ref = ["self", "feature_transform", "sample_transform", "kernel_size", "stride", "in_channel",
           "dim", "ext", "drop", "heinit", "cvinit"]                             # PointNetCls.py:171
    assert list(sig.parameters)[:len(ref)] == ref
    defaults = {k: v.default for k, v in sig.parameters.items() if k != "self"}
    assert defaults["feature_transform"] is False and defaults["sample_transform"] is True
    assert defaults["drop"] == 0.3 and defaults["ext"] is False
    for name in ("STN3d", "STNkd", "PointNetfeat", "PointNetCls", "feature_transform_regularizer"):
        assert hasattr(m, name)


def test_state_dict_keys_match_golden_parameter_set(golden_pointnet):
    """Every parameter the reference exposes (names recorded in the golden file from the reference's
    own named_parameters()) exists here with the same name."""
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    from oracle.gen_golden import POINTNET_CASES
    import oracle
    for case, (kw, B, N, seed) in POINTNET_CASES.items():
        net = PointNetCls(**kw)
        sd = net.state_dict()
        names = {k[len("gsample/"):] for k in golden_pointnet[case] if k.startswith("gsample/")}
        mine = {k for k, _ in net.named_parameters()}
        assert names <= mine
        shapes = oracle.pointnet_oracle.param_shapes(oracle.PointNetSpec(**kw))
        for k, shp in shapes.items():
            assert tuple(sd[k].shape) == tuple(shp), k
        # InstanceNorm buffers and num_batches_tracked are part of the reference's checkpoints
        assert "feat.stn.in5.num_batches_tracked" in sd and "in2.running_var" in sd
        assert "feat.bn3.num_batches_tracked" in sd

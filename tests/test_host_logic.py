"""CPU checks of the host-side logic around the C-ABI: no silent CPU fallback in the new entry points, the
roofline bookkeeping of bench.py (every libpcuda kernel of the profiled step has an algorithmic-work model and a
committed DRAM-traffic figure)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pointcloududa_b200 import ops  # noqa: E402
from pointcloududa_b200._lib import PcudaError  # noqa: E402


def test_fused_sgd_and_bce_refuse_cpu_tensors():
    from pointcloududa_b200.optim import FusedSGD
    p = torch.nn.Parameter(torch.zeros(4, 3))
    with pytest.raises(PcudaError):
        FusedSGD([p], lr=0.1, momentum=0.9)
    with pytest.raises(PcudaError):
        ops.bce_with_logits(torch.zeros(4, 1), 1.0)


def test_kernel_work_table_covers_the_profiled_step():
    import bench
    path = os.path.join(ROOT, "profiles", "traffic_cfg2.json")
    assert os.path.exists(path), "profiles/traffic_cfg2.json is committed evidence (tools/ncu_traffic.py)"
    d = json.load(open(path))
    framework = ("at::", "cutlass", "Memset", "Memcpy", "memset", "memcpy", "nccl")
    ours = [k for k in d["kernels"] if not any(f in k for f in framework)]
    assert len(ours) >= 30
    w = dict(bench.WORKLOADS["cfg2"])
    for k in ours:
        hits = [frag for frag in bench.KERNEL_WORK if frag in k]
        assert hits, f"no algorithmic-work model for kernel {k}"
        flops_f, bytes_f = bench.KERNEL_WORK[hits[0]]
        assert bytes_f(w) > 0 and (flops_f is None or flops_f(w) > 0)
        traffic, src = bench.ncu_traffic("cfg2", "pcuda::" + k)
        assert traffic is not None and traffic >= 0 and "traffic_cfg2.json" in src
        rec = bench.kernel_roofline("pcuda::" + k, w, 10e-6, bench.peaks(), traffic)
        assert rec["bound"] in ("hbm", "tensor", "l2-latency", "latency", "fp32-issue") and rec["frac"] > 0


def test_kernel_roofline_picks_the_limiting_roof():
    import bench
    pk = bench.peaks()
    w = dict(B=32, N=4096, C=5, H=256, W=256, return_prob=True)
    pooled = bench.kernel_roofline("pcuda::tc::ws_kernel<1>", w, 46e-6, pk)
    assert pooled["bound"] == "tensor" and 0.3 < pooled["frac"] < 0.7
    dgrad = bench.kernel_roofline("pcuda::tc::ws_kernel<2>", w, 39e-6, pk)
    assert dgrad["bound"] == "hbm" and dgrad["frac_hbm"] > 5 * dgrad["frac_tensor"]
    small = bench.kernel_roofline("pcuda::pool_sparse_kernel<4, 12, 1>", dict(bench.WORKLOADS["cfg2"]), 20e-6, pk, traffic=1.8e6)
    assert small["bound"] == "l2-latency" and small["traffic_over_algorithmic"] < 0.5


def test_step_input_layout_offsets_are_aligned():
    # the flat layout arithmetic of AdversarialStep (no device needed for the offsets)
    from pointcloududa_b200.step import AdversarialStep
    obj = AdversarialStep.__new__(AdversarialStep)
    obj._input_shapes = {"oS": (8, 4, 224, 224), "vertS": (8, 300, 3), "vertT": (8, 300, 3)}
    offs, total = obj._input_offsets()
    assert all(o % 64 == 0 for o, _, _ in offs.values()) and total % 64 == 0
    flat = torch.zeros(total)
    views = obj.input_layout(flat)
    assert views["vertS"].shape == (8, 300, 3) and views["vertT"].data_ptr() - views["vertS"].data_ptr() == 7232 * 4
    with pytest.raises(ValueError):
        obj.input_layout(torch.zeros(total + 1))


def test_share_forward_context_nests_and_restores():
    """share_forward / ForwardMemo (networks/PointNetCls.py): the memo is thread-local state of the fused ops; the context
    manager installs it, nests, and restores the previous one (also on exceptions).  No CUDA needed."""
    from pointcloududa_b200.networks import PointNetCls as P
    assert getattr(P._TLS, "memo", None) is None
    a, b = P.ForwardMemo(), P.ForwardMemo()
    with P.share_forward(a):
        assert P._TLS.memo is a
        with P.share_forward(b):
            assert P._TLS.memo is b
        assert P._TLS.memo is a
        with P.share_forward(None):           # an inner region that must not share
            assert P._TLS.memo is None
        assert P._TLS.memo is a
        try:
            with P.share_forward(b):
                raise RuntimeError("x")
        except RuntimeError:
            pass
        assert P._TLS.memo is a
    assert P._TLS.memo is None
    a.entries["k"] = 1
    a.clear()
    assert not a.entries
    t = __import__("torch").zeros(3, 4)[:, 1:]
    assert P._tkey(t) == (t.data_ptr(), (3, 3), (4, 1)) and P._tkey(None) is None


def test_roofline_labels_small_launches_as_latency():
    """A launch whose whole roofline time is below a microsecond is labelled 'latency' (bench.kernel_roofline): the cfg2-sized
    FC layer; the same kernel name at a roofline-sized shape keeps its HBM / tensor label."""
    import bench
    pk = {"hbm": 6546.2, "tensor": 1666.9, "source": "test"}
    w = dict(bench.WORKLOADS["cfg2"])
    r = bench.kernel_roofline("pcuda::fc_fwd_kernel", w, 5.8e-6, pk, traffic=9.4e5)
    assert r["bound"] == "latency" and r["roofline_time_us"] < 1.0 and 0 < r["frac"] < 0.1
    big = dict(bench.WORKLOADS["cfg5_rank"])
    r = bench.kernel_roofline("pcuda::tc::ws_kernel<1>", big, 45e-6, pk, traffic=None)
    assert r["bound"] == "tensor" and 0.3 < r["frac"] < 0.6
    r = bench.kernel_roofline("pcuda::entropy_bwd_vec4<1, 5, true, true, false>", big, 170e-6, pk, traffic=None)
    assert r["bound"] == "hbm"
    for name in ("pool_sparse_sorted_kernel", "gram_finish_kernel", "reduce_partials_split_kernel<float>", "sgd_momentum_sum_kernel",
                 "chamfer_bwd_small_kernel", "q_finish_kernel", "pool_q_kernel"):
        assert any(frag in name for frag in bench.KERNEL_WORK), name

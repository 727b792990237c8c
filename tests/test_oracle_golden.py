"""The oracle (CPU restatement) against golden vectors produced by the reference itself
(oracle/gen_golden.py).  This is what pins the oracle; the CUDA parity tests then lean on it."""
import numpy as np
import pytest

import oracle
from oracle.gen_golden import POINTNET_CASES, sample_index
from conftest import max_rel


# ---------------------------------------------------------------- Chamfer -------------------------
def test_chamfer_cases_present(golden_chamfer):
    assert len(golden_chamfer) >= 6


@pytest.mark.parametrize("case", ["near_b3_n64", "near_b2_n300", "uniform_b2_n257",
                                  "identical_dups_b2_n128", "zero_gt_b2_n96", "shifted_b2_n80"])
def test_chamfer_indices_bit_exact(golden_chamfer, case):
    g = golden_chamfer[case]
    o = oracle.chamfer_fwd(g["x"], g["y"])
    # indices: the bit-exact gate
    assert np.array_equal(o["i1"], g["i1"]), f"i1 mismatches: {(o['i1'] != g['i1']).sum()}"
    assert np.array_equal(o["i2"], g["i2"]), f"i2 mismatches: {(o['i2'] != g['i2']).sum()}"
    # distances: torch's CPU sqrt is not correctly rounded (1 ulp off in <1% of elements,
    # SURVEY.md §7); the oracle uses IEEE sqrtf like CUDA does.
    for k in ("d1", "d2"):
        ulp = np.abs(o[k].view(np.int32).astype(np.int64) - g[k].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, (k, ulp.max())
        assert (ulp > 0).mean() < 0.05
    assert abs(float(o["loss"]) - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))


@pytest.mark.parametrize("case", ["near_b3_n64", "near_b2_n300", "shifted_b2_n80", "identical_dups_b2_n128"])
def test_chamfer_pairwise_matrix_bit_exact(golden_chamfer, case):
    """P + 1e-5 of sample 0, bit for bit: pins the fmaf chain and the rounding order."""
    g = golden_chamfer[case]
    P = oracle.pairwise_dist(g["x"][0], g["y"][0])
    q = P + np.float32(0.00001)
    assert q.dtype == np.float32
    assert np.array_equal(q.view(np.int32), g["q1_s0"].view(np.int32))


def test_chamfer_known_answers(golden_chamfer):
    g = golden_chamfer["zero_gt_b2_n96"]
    assert (g["i1"] == 0).all()          # all-zero GT cloud: every x_i picks index 0
    o = oracle.chamfer_fwd(g["x"], g["y"])
    assert (o["i1"] == 0).all()
    g = golden_chamfer["identical_dups_b2_n128"]
    o = oracle.chamfer_fwd(g["x"], g["y"])
    # x == y with y[64:] == y[:64]: the match is never in the upper (duplicate) half unless an
    # even earlier identical point exists
    assert (o["i1"] < 64).all() and (o["i2"] < 64).all()
    assert np.allclose(o["d1"], np.sqrt(np.float32(1e-5)), atol=1e-6)


@pytest.mark.parametrize("case", ["near_b3_n64", "near_b2_n300", "uniform_b2_n257"])
def test_chamfer_gradient_vs_reference_autograd(golden_chamfer, case):
    """The analytic gradient through the reference's indices reproduces its autograd result up to
    the autograd path's own cancellation noise (SURVEY.md §7 'Gradient tolerance')."""
    g = golden_chamfer[case]
    gx = oracle.chamfer_bwd64(g["x"], g["y"], g["d1"], g["i1"], g["d2"], g["i2"])
    assert max_rel(g["grad_x"], gx) < 5e-3


# ---------------------------------------------------------------- entropy -------------------------
ENT_CASES = ["sigmoid_c4", "softmax_c5_norm", "sigmoid_c5_norm", "softmax_c4_confident",
             "softmax_zero", "sigmoid_zero", "sigmoid_saturated"]


@pytest.mark.parametrize("case", ENT_CASES)
def test_entropy_forward(golden_entropy, case):
    g = golden_entropy[case]
    m, p, mean = oracle.entropy_fwd(g["z"], str(g["act"]), bool(g["norm"]), want_p=True, want_mean=True)
    # tolerance from SURVEY.md §7: |a-b| <= 1e-5*|b| + 5e-7
    assert np.all(np.abs(m - g["m"]) <= 1e-5 * np.abs(g["m"]) + 5e-7)
    assert np.all(np.abs(p - g["p"]) <= 1e-5 * np.abs(g["p"]) + 1e-9)
    assert abs(float(mean) - float(g["mean"])) <= 1e-5 * abs(float(g["mean"])) + 1e-7


def test_entropy_known_answers(golden_entropy):
    import math
    g = golden_entropy["softmax_zero"]
    closed = -(1 / 5) * math.log(1 / 5 + 1e-7) / math.log(5)
    assert np.allclose(g["m"], closed, rtol=1e-6)
    g = golden_entropy["sigmoid_zero"]
    assert np.allclose(g["m"], -0.5 * math.log(0.5 + 1e-7), rtol=1e-6)
    g = golden_entropy["sigmoid_saturated"]
    assert g["m"].min() < 0 and g["m"].min() > -2e-7   # tiny negative entries for saturated logits


@pytest.mark.parametrize("case", ENT_CASES)
def test_entropy_backward(golden_entropy, case):
    g = golden_entropy[case]
    gz = oracle.entropy_bwd64(g["z"], g["grad_m"], g["grad_p"], float(g["grad_mean"]),
                              str(g["act"]), bool(g["norm"]))
    # max-norm relative 1e-5 plus an absolute floor: for saturated logits fp32 rounds 1-p to 0
    # where fp64 keeps ~1e-9 (the reference itself returns 0 there)
    assert np.abs(g["grad_z"] - gz).max() <= 1e-5 * np.abs(gz).max() + 1e-8


# ---------------------------------------------------------------- PointNet ------------------------
# (forward, gradient) max-norm tolerances = ~4x the measured distance between the reference's own
# fp32 result and the float64 oracle (printed by the comment in oracle/gen_golden.py's history):
# the network is a chain of tiny-batch BatchNorms, so its conditioning varies by case.
POINTNET_TOL = {
    "default_b4_n64": (2e-5, 3e-4),
    "default_b6_n300": (2e-5, 1e-4),
    "ft_b5_n48": (2e-4, 2e-3),
    "ft_ext_b4_n40": (5e-4, 1e-2),
    "default_b2_n50": (5e-3, None),     # batch of 2: BN output is +-1; gradients are chaotic
    "nostn_b4_n32": (2e-5, 1e-4),
}

@pytest.mark.parametrize("case", list(POINTNET_CASES))
def test_pointnet_forward_backward(golden_pointnet, case):
    g = golden_pointnet[case]
    kw, B, N, seed = POINTNET_CASES[case]
    spec = oracle.PointNetSpec(**kw)
    params = oracle.make_params(spec, seed)
    x = np.ascontiguousarray(g["pts"].transpose(0, 2, 1))
    # The oracle is evaluated in float64 and the reference's float32 result must sit within ITS OWN
    # rounding noise of it.  That noise depends on conditioning: BatchNorm over a batch of 2-3 rows
    # in the FC layers amplifies fp32 rounding ~1e3x, a batch of 4 does not (measured:
    # reference-vs-fp64 5e-6 at B=4, 2e-5..1.2e-4 at B=2..3).
    tol, gtol = POINTNET_TOL[case]
    logit, trans, trans_feat, cache = oracle.pointnet_forward(params, x, spec, dtype=np.float64)
    assert max_rel(logit, g["logit"]) < tol
    if trans is not None:
        assert max_rel(trans, g["trans"]) < tol
    if trans_feat is not None:
        assert max_rel(trans_feat.sum(axis=(1, 2)), g["trans_feat_sum"]) < tol
    # loss = BCE-with-logits vs 1 (train_mscmrseg.py:233): dL/dlogit = (sigmoid(l) - 1)/B
    dlogit = ((1.0 / (1.0 + np.exp(-logit.astype(np.float64)))) - 1.0) / logit.shape[0]
    if gtol is None:
        return
    dx, grads = oracle.pointnet_backward(dlogit, cache)
    assert max_rel(dx, g["grad_x"]) < gtol
    checked = 0
    for k, gs in g.items():
        if not k.startswith("gsample/"):
            continue
        name = k[len("gsample/"):]
        mine = grads[name].reshape(-1)
        norm_ref = float(g[f"gnorm/{name}"])
        idx = sample_index(mine.size, name)
        scale = max(np.abs(gs).max(), norm_ref / np.sqrt(mine.size), 1e-12)
        # conv/fc biases feeding a train-mode BN have a mathematically zero gradient; both sides
        # hold rounding noise there, compare on the absolute scale of the layer's weight grads
        leaf = name.rsplit(".", 2)[-2]
        if name.endswith(".bias") and (leaf.startswith("conv") or leaf in ("fc1", "fc2")):
            # mathematically zero (train-mode BN cancels the preceding bias): exact in fp64, fp32
            # rounding noise in the reference
            assert np.abs(mine).max() < 1e-9 and np.abs(gs).max() < 1e-3, name
        else:
            # + absolute floor: some gradients are mathematically zero (e.g. feat.bn3.bias, whose
            # per-channel shift is removed by the head's train-mode bn1) and hold only noise
            assert np.abs(mine[idx] - gs).max() <= gtol * scale + gtol * 1e-3, name
            mynorm = np.sqrt((mine.astype(np.float64) ** 2).sum())
            assert abs(mynorm - norm_ref) <= gtol * norm_ref + gtol * 1e-3 * np.sqrt(mine.size), name
        checked += 1
    assert checked >= 20


def test_pointnet_running_stats(golden_pointnet):
    """BatchNorm buffers after one forward: 0.9*init + 0.1*batch stat (unbiased var)."""
    case = "default_b4_n64"
    g = golden_pointnet[case]
    kw, B, N, seed = POINTNET_CASES[case]
    spec = oracle.PointNetSpec(**kw)
    params = oracle.make_params(spec, seed)
    x = np.ascontiguousarray(g["pts"].transpose(0, 2, 1))
    _, _, _, cache = oracle.pointnet_forward(params, x, spec, dtype=np.float64)
    mean, var = cache["post"]["stats"][-1]      # feat.bn3
    M = B * N
    assert max_rel(0.1 * mean, g["buf/feat.bn3.running_mean"]) < 1e-4
    assert max_rel(0.9 + 0.1 * var * M / (M - 1), g["buf/feat.bn3.running_var"]) < 1e-4


def test_bf16_operand_rounding_matches_torch():
    """oracle.bf16_round (used to restate the tensor-core path's operand rounding) == torch's
    round-to-nearest-even float32 -> bfloat16 conversion, including ties and denormal-range values."""
    import torch
    rs = np.random.RandomState(0)
    a = np.concatenate([rs.standard_normal(50000).astype(np.float32) * np.float32(3.0),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, -1.00390625, 3.4e38, 1e-39, 65280.0], np.float32)])
    want = torch.from_numpy(a).bfloat16().float().numpy()
    assert np.array_equal(oracle.bf16_round(a), want)
    assert np.array_equal(oracle.bf16_round(a.astype(np.float64)), want.astype(np.float64))
    assert oracle.bf16_layer(64) and oracle.bf16_layer(512) and not oracle.bf16_layer(3) and not oracle.bf16_layer(1024)


def test_pointnet_oracle_bf16_operands_only_touch_wide_layers():
    spec = oracle.PointNetSpec()
    params = oracle.make_params(spec, 1)
    x = np.random.RandomState(2).rand(4, 3, 40).astype(np.float32)
    layers = [dict(weight=params["feat.conv1.weight"], bias=params["feat.conv1.bias"], gamma=params["feat.bn1.weight"],
                   beta=params["feat.bn1.bias"], relu=True)]
    a, _ = oracle.stack_forward(x, layers, False, np.float64)
    b, _ = oracle.stack_forward(x, layers, False, np.float64, bf16_operands=True)
    assert np.array_equal(a, b)                       # K = 3 layer stays full precision
    l0 = oracle.pointnet_forward(params, x, spec, dtype=np.float64)[0]
    l1 = oracle.pointnet_forward(params, x, spec, dtype=np.float64, bf16_operands=True)[0]
    assert not np.array_equal(l0, l1) and np.abs(l0 - l1).max() < 0.5 * np.abs(l0).max()

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


# ------------------------------------------------- torch restatement of the step (oracle/torch_step.py) ----
def _torch_sd(spec, params, dtype):
    """state_dict-shaped dict of torch tensors (parameters from `params`, fresh BatchNorm buffers)."""
    import torch
    from oracle import torch_step  # noqa: F401
    sd = {}
    for k, v in params.items():
        sd[k] = torch.from_numpy(v.copy()).to(dtype)
        mod = k.rsplit(".", 1)[0]
        if mod.rsplit(".", 1)[-1].startswith("bn") and k.endswith(".weight"):
            sd[mod + ".running_mean"] = torch.zeros(v.shape, dtype=dtype)
            sd[mod + ".running_var"] = torch.ones(v.shape, dtype=dtype)
            sd[mod + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return sd


@pytest.mark.parametrize("case", list(POINTNET_CASES))
def test_torch_restatement_vs_reference_golden(golden_pointnet, case):
    """oracle/torch_step.D4 (the oracle the GPU step-parity tests lean on) reproduces the reference's own logits,
    transforms, input gradient, sampled parameter gradients and BatchNorm buffers."""
    import torch
    import torch.nn.functional as F
    from oracle import torch_step
    g = golden_pointnet[case]
    kw, B, N, seed = POINTNET_CASES[case]
    spec = oracle.PointNetSpec(**kw)
    tol, gtol = POINTNET_TOL[case]
    sd = _torch_sd(spec, oracle.make_params(spec, seed), torch.float64)
    keys = torch_step.trainable_keys(sd)
    leaves = {k: (v.requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    x = torch.from_numpy(g["pts"]).double().transpose(2, 1).requires_grad_(True)
    net = torch_step.D4(leaves, feature_transform=spec.feature_transform, ext=spec.ext, sample_transform=spec.sample_transform)
    logit, trans, trans_feat = net(x)
    assert max_rel(logit.detach().numpy(), g["logit"]) < tol
    if trans is not None:
        assert max_rel(trans.detach().numpy(), g["trans"]) < tol
    if trans_feat is not None:
        assert max_rel(trans_feat.detach().sum(dim=(1, 2)).numpy(), g["trans_feat_sum"]) < tol
    for k, v in sd.items():
        if f"buf/{k}" in g:
            assert max_rel(v.detach().numpy(), g[f"buf/{k}"]) < 1e-4, k
    if gtol is None:
        return
    loss = F.binary_cross_entropy_with_logits(logit, torch.ones_like(logit))
    used = [k for k in keys if spec.sample_transform or not k.startswith("feat.stn.")]
    grads = torch.autograd.grad(loss, [x] + [leaves[k] for k in used])
    assert max_rel(grads[0].numpy(), g["grad_x"]) < gtol
    for k, gr in zip(used, grads[1:]):
        gs = g[f"gsample/{k}"]
        mine = gr.numpy().reshape(-1)
        idx = sample_index(mine.size, k)
        scale = max(np.abs(gs).max(), float(g[f"gnorm/{k}"]) / np.sqrt(mine.size), 1e-12)
        leaf = k.rsplit(".", 2)[-2]
        if k.endswith(".bias") and (leaf.startswith("conv") or leaf in ("fc1", "fc2")):
            assert np.abs(mine).max() < 1e-9, k
        else:
            assert np.abs(mine[idx] - gs).max() <= gtol * scale + gtol * 1e-3, k


@pytest.mark.parametrize("kw", [dict(), dict(feature_transform=True, ext=True)])
def test_torch_restatement_bf16_equals_numpy_oracle(kw):
    """The two same-rounding oracles (numpy pointnet_forward/backward with bf16_operands, torch D4 with
    bf16_operands) agree in float64: logits, input gradient, every parameter gradient."""
    import torch
    import torch.nn.functional as F
    from oracle import torch_step
    spec = oracle.PointNetSpec(**kw)
    params = oracle.make_params(spec, 5)
    rs = np.random.RandomState(6)
    pts = (rs.rand(4, 48, 3) * rs.uniform(0.3, 1.0, (4, 1, 3))).astype(np.float32)
    xo = np.ascontiguousarray(pts.transpose(0, 2, 1))
    lo, _, _, cache = oracle.pointnet_forward(params, xo, spec, dtype=np.float64, bf16_operands=True)
    dlogit = ((1.0 / (1.0 + np.exp(-lo))) - 1.0) / lo.shape[0]
    dx, grads = oracle.pointnet_backward(dlogit, cache)
    sd = _torch_sd(spec, params, torch.float64)
    keys = torch_step.trainable_keys(sd)
    leaves = {k: (v.requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    x = torch.from_numpy(pts).double().transpose(2, 1).requires_grad_(True)
    logit = torch_step.D4(leaves, bf16_operands=True, **kw)(x)[0]
    assert max_rel(logit.detach().numpy(), lo) < 1e-9
    loss = F.binary_cross_entropy_with_logits(logit, torch.ones_like(logit))
    gr = torch.autograd.grad(loss, [x] + [leaves[k] for k in keys])
    assert max_rel(gr[0].numpy(), dx) < 1e-7
    for k, t in zip(keys, gr[1:]):
        ref = grads[k].reshape(t.shape)
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(t.numpy() - ref).max() <= 1e-6 * scale + 1e-12, k


def test_torch_restatement_step_runs_and_moves_parameters():
    """One restated step on the CPU: finite results, BatchNorm buffers advanced three times, SGD applied."""
    import torch
    from oracle import torch_step
    from pointcloududa_b200.synth import synth_clouds, synth_logits
    spec = oracle.PointNetSpec()
    sd = _torch_sd(spec, oracle.make_params(spec, 2), torch.float64)
    before = {k: v.clone() for k, v in sd.items()}
    B, N = 4, 64
    xS, yA = synth_clouds(B, N, seed=1, mode="near")
    xT, yB = synth_clouds(B, N, seed=2, mode="uniform")
    inputs = dict(oS=synth_logits(B, 4, 16, 16, seed=3), oT=synth_logits(B, 4, 16, 16, seed=4),
                  grad_mapT=synth_logits(B, 4, 16, 16, seed=5, std=1.0), vertS=xS, vertT=xT, vertexA=yA, vertexB=yB)
    inputs = {k: v.double() for k, v in inputs.items()}
    mom = {}
    out = torch_step.adversarial_step(sd, inputs, mom)
    assert torch.isfinite(out["result"]).all() and len(out["result"]) == len(torch_step.RESULT_FIELDS)
    assert int(sd["feat.bn3.num_batches_tracked"]) == 3
    assert not torch.equal(sd["fc3.weight"], before["fc3.weight"])
    w = "feat.conv3.weight"
    want = before[w] - 2.5e-5 * (out["grads"][w] + 0.0005 * before[w])
    assert torch.allclose(sd[w], want, rtol=0, atol=1e-15)
    assert out["grad_vertS"].shape == (B, N, 3) and out["grad_oT"].shape == inputs["oT"].shape


def test_reference_step_equals_torch_restatement():
    """The reference's own code run as one step (baseline/ref_step.py over baseline/_ref, the arm bench.py times)
    and the restated step agree: this pins oracle/torch_step.adversarial_step — every result field, the three input
    gradients, the parameters and BatchNorm buffers after SGD — to the reference itself.  Needs baseline/_ref
    (python baseline/make_ref.py in a container that has /root/reference)."""
    import torch
    from baseline import ref_step
    from oracle import torch_step
    from bench import WORKLOADS
    if not ref_step.available():
        pytest.skip("baseline/_ref not installed")
    torch.set_num_threads(2)
    w = dict(WORKLOADS["cfg2"]); w.update(B=6, H=32, W=32, N=64)
    host = torch_step.conditioned_inputs(w, 11)
    ref = ref_step.ReferenceStep(w, "cpu", seed=0)
    for m in ref.model_dis4.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    sd = {k: (v.detach().clone().double() if v.is_floating_point() else v.clone()) for k, v in ref.model_dis4.state_dict().items()}
    mom = {}
    start = {k: v.clone() for k, v in sd.items()}
    inputs = {k: v.double() for k, v in host.items()}
    for it in range(2):
        res = ref.run(host)
        out = torch_step.adversarial_step(sd, inputs, mom, activation="sigmoid", normalize=False)
        assert max_rel(res.numpy(), out["result"].numpy()) < 2e-5, it
        assert max_rel(ref.grad_vertS.numpy(), out["grad_vertS"].numpy()) < 5e-3       # autograd's own cancellation noise
        assert max_rel(ref.grad_oT.numpy(), out["grad_oT"].numpy()) < 1e-5
        assert max_rel(ref.grad_vertT.numpy(), out["grad_vertT"].numpy()) < 1e-3
        now = ref.model_dis4.state_dict()
        for k, v in sd.items():
            leaf = k.rsplit(".", 2)[-2] if k.count(".") else k
            if leaf.startswith("in"):
                continue
            if k.endswith("num_batches_tracked"):
                assert int(now[k]) == int(v) == 3 * (it + 1), k
            elif "running" in k:
                assert max_rel(now[k].numpy(), v.numpy()) < 1e-4, k
            else:
                # parameters move by lr * grad ~ 1e-5: compared on the scale of the accumulated update (gradient
                # accuracy of the fp32 reference through batch-of-4 BatchNorms) plus fp32 resolution of the value
                upd = (v - start[k]).abs().max().item()
                err = np.abs(now[k].double().numpy() - v.numpy()).max()
                assert err <= 5e-3 * upd + 2e-7 * v.abs().max().item() + 1e-10, (k, err, upd)   # 1e-10: mathematically zero gradients


# ---------------------------------------------------------------- farthest-point sampling ------------
@pytest.mark.parametrize("case", ["grid_v2000_k300", "uniform_v5000_k300", "dim2_v700_k64", "dups_v512_k40", "k_exceeds_v"])
def test_fps_oracle_vs_reference_golden(golden_fps, case):
    """oracle.fps (graipher with the start index explicit) reproduces the reference's own selection bit for bit."""
    g = golden_fps[case]
    far, idx = oracle.fps(g["pts"], g["far"].shape[0], int(g["start"]))
    assert np.array_equal(far, g["far"])
    assert idx[0] == int(g["start"])
    if case == "k_exceeds_v":           # once every distance is 0 np.argmax returns index 0 for good
        assert (idx[g["pts"].shape[0]:] == 0).all()

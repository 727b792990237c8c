"""GPU parity of the BENCHED paths against oracles that share nothing with them:

* the tcgen05 (bf16-operand) shared-MLP backward and the whole bf16 discriminator's gradients against the
  numpy oracle evaluated in float64 with the SAME operand rounding (`bf16_operands=True`) — north_star's
  1e-2 bar for the bf16 MLP, max-norm;
* one whole `AdversarialStep.run()` (fp32 and bf16, eager and CUDA-graph replay) against the plain-torch
  restatement of train_mscmrseg.py:183-330 (`oracle/torch_step.py`, pinned to the reference's golden outputs
  by tests/test_oracle_golden.py) evaluated in float64 on the same GPU: all 8 result fields, the gradients into
  both clouds and the target logits, D4's summed parameter gradients, its parameters and BatchNorm buffers
  after the SGD step.

A max-pool arg-max or a ReLU whose two candidates are closer than the working precision can resolve may
legitimately switch between two correct evaluations and moves a whole gradient column; stack-level tests mask
such (cloud, channel) pairs out of the upstream gradient (like tests/test_gpu_parity.py::_run_stack_instance),
whole-network tests bound the fraction of entries outside the bar instead.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from conftest import max_rel
from oracle import torch_step
from pointcloududa_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_TOL = 1e-2          # BASELINE.json north_star: "within ... 1e-2 (bf16 MLP)"
AMBIGUOUS = 2e-3         # relative top-1 / top-2 gap below which a max-pool selection is not comparable across evaluations


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


# ------------------------------------------------------------------------------------------------------
# tcgen05 shared-MLP backward vs the same-rounding float64 oracle
ORACLE_STACKS = [
    # chans, relus, B, N
    ([3, 64, 128, 1024], [True, True, True], 4, 64),        # STN3d trunk, partial tile
    ([3, 64, 128, 1024], [True, True, False], 8, 300),      # main trunk, cfg-1/2 (benched)
    ([3, 64, 128, 1024], [True, True, True], 8, 300),       # STN3d trunk, cfg-1/2 (benched)
    ([3, 64, 128, 1024], [True, True, False], 4, 1024),     # cfg-3 per-rank shard
    ([3, 64, 128, 1024], [True, True, False], 32, 4096),    # cfg-5 per-rank shard (benched)
    ([64, 64, 128, 1024], [True, True, True], 3, 70),       # STNkd trunk
    ([64, 128, 1024], [True, False], 4, 100),               # feature-transform post stack
    ([3, 8, 64, 128, 256, 512, 1024], [True, True, True, True, False, True], 3, 40),   # ext trunk (K up to 512)
]


def _make_stack(chans, seed):
    import torch.nn as nn
    rs = np.random.RandomState(seed)
    torch.manual_seed(seed)
    convs, bns = [], []
    for l in range(len(chans) - 1):
        conv = nn.Conv1d(chans[l], chans[l + 1], 1).to(DEV)
        bn = nn.BatchNorm1d(chans[l + 1]).to(DEV)
        with torch.no_grad():
            bn.weight.copy_(torch.from_numpy(1.0 + 0.5 * rs.standard_normal(chans[l + 1])).float())
            bn.bias.copy_(torch.from_numpy(0.2 * rs.standard_normal(chans[l + 1])).float())
        convs.append(conv); bns.append(bn)
    return convs, bns, rs


def stack_errors(chans, relus, B, N, seed):
    """max-norm and L2 relative errors of the tensor-core stack (forward value, dx, every dW / dgamma / dbeta)
    against oracle.stack_forward / stack_backward in float64 with bf16 operand rounding."""
    from pointcloududa_b200.networks.PointNetCls import shared_mlp
    convs, bns, rs = _make_stack(chans, seed)
    scale = rs.uniform(0.3, 1.0, size=(B, 1, chans[0]))
    pts = (rs.rand(B, N, chans[0]) * scale).astype(np.float32)
    layers = [dict(weight=c.weight.detach().cpu().numpy(), bias=c.bias.detach().cpu().numpy(),
                   gamma=b.weight.detach().cpu().numpy(), beta=b.bias.detach().cpu().numpy(), relu=r)
              for c, b, r in zip(convs, bns, relus)]
    xo = np.ascontiguousarray(pts.transpose(0, 2, 1))
    out_o, cache = oracle.stack_forward(xo, layers, True, np.float64, bf16_operands=True)
    # pooled tensor of the oracle: (cloud, channel) pairs whose two largest values are within the fp32
    # accumulation noise of each other may select another point on the GPU; mask them out of the upstream gradient
    a_last = cache["caches"][-1][4] if not relus[-1] else np.maximum(cache["caches"][-1][4], 0)      # [C, B*N]
    a3 = a_last.reshape(a_last.shape[0], B, N)
    if N > 1:
        part = np.partition(a3, N - 2, axis=2)
        top1, top2 = part[:, :, N - 1], part[:, :, N - 2]
        # the GPU's pooled values sit within ~5e-4 of the oracle's (a handful of activations straddle a bf16 rounding
        # boundary between the fp32 and the float64 evaluation, and BatchNorm amplifies that): pairs closer than
        # 2e-3 are ambiguous between two correct evaluations
        ambiguous = ((top1 - top2) < AMBIGUOUS * (1.0 + np.abs(top1))).T          # [B, C]
    else:
        ambiguous = np.zeros((B, a3.shape[0]), bool)
    w = rs.standard_normal(out_o.shape)
    w[ambiguous] = 0.0
    dx_o, grads_o = oracle.stack_backward(w, cache)

    x = torch.from_numpy(pts).to(DEV).transpose(2, 1).requires_grad_(True)
    out = shared_mlp(x, convs, bns, relus, pool=True, precision="bf16")
    (out * torch.from_numpy(w).float().to(DEV)).sum().backward()
    torch.cuda.synchronize()
    errs = {}

    def put(name, mine, ref, floor=0.0):
        mine = np.asarray(mine, np.float64).reshape(ref.shape)
        sm = max(np.abs(ref).max(), floor, 1e-30)
        sn = max(np.linalg.norm(ref), floor * np.sqrt(ref.size), 1e-30)
        d = np.abs(mine - ref)
        errs[name] = (float(d.max() / sm), float(np.linalg.norm(mine - ref) / sn), float((d > BF16_TOL * sm).mean()))

    put("out", out.detach().cpu().numpy(), out_o)
    put("dx", x.grad.cpu().numpy(), dx_o)
    for l, (c, b, g) in enumerate(zip(convs, bns, grads_o)):
        gw = np.abs(g["weight"]).max()
        put(f"dW{l}", c.weight.grad.cpu().numpy(), g["weight"])
        # dgamma / dbeta can be mathematically zero (a BatchNorm without ReLU feeding another train-mode
        # BatchNorm): compared on the scale of the layer's weight gradient
        put(f"dgamma{l}", b.weight.grad.cpu().numpy(), g["gamma"], floor=1e-2 * gw)
        put(f"dbeta{l}", b.bias.grad.cpu().numpy(), g["beta"], floor=1e-2 * gw)
    errs["masked_pairs"] = int(ambiguous.sum())
    return errs


@pytest.mark.parametrize("chans,relus,B,N", ORACLE_STACKS)
def test_tensor_core_stack_backward_vs_same_rounding_oracle(chans, relus, B, N):
    """Op-level bar of the north star for the bf16 MLP, against an oracle that shares nothing with the kernels:
    every gradient of the tensor-core shared MLP within 1e-2 (L2) of the float64 oracle with the same operand
    rounding; max-norm within 1e-2 for all but <= 0.1 % of the entries and never beyond 5e-2.  The stragglers
    are ReLU units whose input is within bf16 noise of zero: `tools/bf16_err_probe.py` shows the same figures
    with every backward piece on the FP32 kernels, i.e. they belong to the bf16 forward, not to the tcgen05
    backward.  BatchNorm over fewer than 1024 points amplifies the operand rounding itself: 2e-2 there."""
    errs = stack_errors(chans, relus, B, N, seed=1000 + 7 * B + N)
    masked = errs.pop("masked_pairs")
    print(f"stack {chans} B={B} N={N}: masked pairs {masked}/{B * chans[-1]}; " +
          ", ".join(f"{k} {m:.2e}/{n:.2e}/{f:.1e}" for k, (m, n, f) in errs.items()))
    assert masked < 0.35 * B * chans[-1]
    tol = BF16_TOL if B * N >= 1024 else 2 * BF16_TOL
    for k, (emax, el2, frac) in errs.items():
        assert el2 <= tol, (k, emax, el2, frac)
        assert emax <= 5 * tol, (k, emax, el2, frac)
        assert frac <= 1e-3 or emax <= tol, (k, emax, el2, frac)


# ------------------------------------------------------------------------------------------------------
# whole bf16 discriminator: logits and every gradient vs the same-rounding float64 oracle
def _load_params(net, params):
    sd = net.state_dict()
    for k, v in params.items():
        sd[k] = torch.from_numpy(np.asarray(v).copy())
    net.load_state_dict(sd)


def _oracle_d4(params_sd, pts, kw, dtype):
    """Logits and every gradient of BCE(D4(pts), 1) by the torch restatement with bf16 operand rounding, evaluated in
    `dtype` on the GPU (pinned to the numpy oracle and to the reference's golden vectors by the CPU suite)."""
    sd = {k: (v.detach().to(dtype).clone() if v.is_floating_point() else v.clone()) for k, v in params_sd.items()}
    keys = torch_step.trainable_keys(sd)
    leaves = {k: (v.requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    xr = pts.to(dtype).transpose(2, 1).requires_grad_(True)
    lo = torch_step.D4(leaves, bf16_operands=True, **kw)(xr)[0]
    loss = F.binary_cross_entropy_with_logits(lo, torch.ones_like(lo))
    gr = torch.autograd.grad(loss, [xr] + [leaves[k] for k in keys])
    out = {"logit": lo.detach().double().cpu().numpy(), "dx": gr[0].double().cpu().numpy()}
    out.update({k: g.double().cpu().numpy() for k, g in zip(keys, gr[1:])})
    return out


def d4_errors(kw, B, N, seed):
    """L2 errors of the bf16 discriminator against the float64 same-rounding oracle (`err`), next to the distance between
    a float32 and a float64 evaluation of that SAME oracle arithmetic (`floor`): what the arithmetic itself leaves
    undetermined once a few activations straddle a bf16 rounding boundary and the batch-of-B BatchNorms amplify it."""
    from pointcloududa_b200.networks.PointNetCls import PointNetCls
    spec = oracle.PointNetSpec(**kw)
    params = oracle.make_params(spec, seed)
    rs = np.random.RandomState(seed + 1)
    pts_np = (rs.rand(B, N, 3) * rs.uniform(0.25, 1.0, (B, 1, 3)) + rs.uniform(0.0, 0.3, (B, 1, 3))).astype(np.float32)
    net = PointNetCls(drop=0.0, precision="bf16", **kw)
    _load_params(net, params)
    net.to(DEV).train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    pts = torch.from_numpy(pts_np).to(DEV)
    x = pts.transpose(2, 1).requires_grad_(True)
    logit = net(x)[0]
    F.binary_cross_entropy_with_logits(logit, torch.ones_like(logit)).backward()
    torch.cuda.synchronize()
    mine = {k: p.grad.detach().double().cpu().numpy() for k, p in net.named_parameters() if p.grad is not None}
    mine["dx"] = x.grad.double().cpu().numpy()
    mine["logit"] = logit.detach().double().cpu().numpy()
    o64 = _oracle_d4(sd0, pts, kw, torch.float64)
    o32 = _oracle_d4(sd0, pts, kw, torch.float32)
    res = {}
    for k, ref in o64.items():
        leaf = k.rsplit(".", 2)[-2] if "." in k else ""
        if k.endswith(".bias") and (leaf.startswith("conv") or leaf in ("fc1", "fc2")):
            continue                      # mathematically zero (train-mode BN cancels the preceding bias)
        if k not in mine or np.abs(ref).max() < 1e-12:
            continue
        den = max(np.linalg.norm(ref), 1e-30)
        res[k] = (float(np.linalg.norm(mine[k].reshape(ref.shape) - ref) / den),
                  float(np.linalg.norm(o32[k].reshape(ref.shape) - ref) / den))
    return res


D4_CASES = [
    (dict(), 8, 300),                                   # cfg-1/2 (benched)
    (dict(), 4, 1024),                                  # cfg-3 per-rank shard
    (dict(), 32, 1024),                                 # cfg-3 full batch
    (dict(feature_transform=True), 8, 300),
    (dict(feature_transform=True, ext=True), 8, 300),   # MM-WHS -ft -extd4
    (dict(), 32, 4096),                                 # cfg-5 per-rank shard (benched)
]


@pytest.mark.parametrize("kw,B,N", D4_CASES)
def test_bf16_discriminator_gradients_vs_same_rounding_oracle(kw, B, N):
    """Logits and every gradient of the whole bf16 (tcgen05) discriminator against the float64 oracle with the same
    operand rounding.  The network is a chain of train-mode BatchNorms over 4-32 clouds: one activation landing on
    the other side of a bf16 rounding boundary (0.4 % of its value) is amplified to percents of a gradient, so a
    float32 and a float64 evaluation of the ORACLE ITSELF differ by that much (`floor`, printed) -- measured
    here: 3-35 % of the gradient norm at these shapes.  The bar is therefore 1e-2 or 4x that floor, whichever is
    larger: the kernels must be as close to the oracle as its own arithmetic is determined; the 1e-2 op-level bar is
    enforced on the shared-MLP stacks above, where no FC-head BatchNorm amplifies it."""
    res = d4_errors(kw, B, N, seed=40 + B + N)
    print(f"D4 {kw} B={B} N={N} (err/floor): " + "; ".join(f"{k} {e:.1e}/{f:.1e}" for k, (e, f) in res.items()))
    worst = max(f for _, f in res.values())
    for k, (err, floor) in res.items():
        assert err <= max(BF16_TOL, 4.0 * floor, 2.0 * worst if k != "logit" else 0.0), (k, err, floor, worst)


# ------------------------------------------------------------------------------------------------------
# one whole adversarial step vs the torch restatement of train_mscmrseg.py:183-330
STEP_SHAPES = {
    "cfg2_small_maps": dict(B=8, C=4, H=64, W=64, N=300, activation="sigmoid", normalize=False, return_prob=False),
    "mmwhs_shard": dict(B=4, C=5, H=32, W=32, N=1024, activation="softmax", normalize=True, return_prob=True),
}


def _step_inputs(w, seed):
    return torch_step.conditioned_inputs(w, seed)


def _ref_state(step, dtype=torch.float64):
    return {k: v.detach().clone().to(dtype if v.is_floating_point() else v.dtype) for k, v in step.d4.state_dict().items()}


def _l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _is_bn_buffer(k):
    leaf = k.rsplit(".", 2)[-2] if k.count(".") else k
    return not leaf.startswith("in")


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", list(STEP_SHAPES))
def test_adversarial_step_vs_torch_restatement(shape, precision, graph):
    """Two consecutive `AdversarialStep.run()` (eager launches or CUDA-graph replay) against the torch restatement of
    train_mscmrseg.py:183-330 evaluated in float64 on the same GPU (with the tensor-core path's operand rounding for
    precision="bf16").  Bars: the discriminator-free quantities at the north star's 1e-5 (5e-5 for the Chamfer gradient
    against float64: the reference's own fp32 `P = rx + ry - 2 zz` carries 1e-4 relative cancellation noise for
    near-coincident points, reproduced bit for bit by the kernel); everything that passes through D4 at 2e-4 (fp32)
    / 1e-2 (bf16) or 4x the distance between a float32 and the float64 evaluation of the restatement itself,
    whichever is larger (see test_bf16_discriminator_gradients_vs_same_rounding_oracle)."""
    from pointcloududa_b200.step import RESULT_FIELDS, AdversarialStep, StepConfig
    assert RESULT_FIELDS == torch_step.RESULT_FIELDS
    w = STEP_SHAPES[shape]
    lr = 2.5e-3     # 100x the reference's lr_dis so that the parameter update is visible above fp32 resolution
    cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"],
                     normalize=w["normalize"], return_prob=w["return_prob"], precision=precision, lr_dis=lr)
    dev = torch.device(DEV)
    st = AdversarialStep(cfg, dev, seed=3)
    for m in st.d4.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    host = _step_inputs(w, seed=77)
    st.load_inputs(host, non_blocking=False)
    if graph:
        st.capture(warmup=2)           # must leave parameters, momentum, BatchNorm buffers and RNG untouched
    bf16 = precision == "bf16"
    base = BF16_TOL if bf16 else 2e-4
    kw = dict(activation=w["activation"], normalize=w["normalize"], dr=cfg.dr, lr=lr, bf16_operands=bf16)
    for it in range(2):                          # second step: momentum buffer, moved parameters, BN buffers reused
        # each step is checked as a map from the product's own state before it (parameters, BatchNorm buffers, SGD
        # momentum): with train-mode BatchNorm over a handful of clouds two trajectories that differ by rounding
        # drift apart chaotically, which would measure the network, not the implementation
        torch.cuda.synchronize()
        sd = {dt: _ref_state(st, dt) for dt in (torch.float64, torch.float32)}
        p_before = {k: v.clone() for k, v in sd[torch.float64].items()}
        mom, off = {torch.float64: {}, torch.float32: {}}, 0
        for k, p_ in zip(st._pnames, st.bucket.params):
            for dt in mom:
                mom[dt][k] = st.opt.momentum_buffer[off:off + p_.numel()].view_as(p_).to(dt).clone()
            off += p_.numel()
        got = st.run().clone().double().cpu().numpy()
        torch.cuda.synchronize()
        out = {dt: torch_step.adversarial_step(sd[dt], {k: v.to(dt).to(dev) for k, v in host.items()}, mom[dt], **kw)
               for dt in (torch.float64, torch.float32)}
        o64, o32 = out[torch.float64], out[torch.float32]
        ref, ref32 = o64["result"].cpu().numpy(), o32["result"].double().cpu().numpy()
        report = {f: (got[i], ref[i], ref32[i]) for i, f in enumerate(RESULT_FIELDS)}
        for i, f in enumerate(RESULT_FIELDS):
            if f.startswith("d4_acc"):
                assert abs(got[i] - ref[i]) <= 1.0 / w["B"] + 1e-6, (it, f, report)     # one logit at the 0.5 boundary at most
            elif f in ("chamfer_source", "chamfer_target", "entropy_mean_target"):
                assert abs(got[i] - ref[i]) <= 1e-5 * abs(ref[i]) + 1e-7, (it, f, report)
            else:
                assert abs(got[i] - ref[i]) <= max(base * abs(ref[i]), 4 * abs(ref32[i] - ref[i])), (it, f, report)
        assert max_rel(st.grad_vertS.cpu().numpy(), o64["grad_vertS"].cpu().numpy()) < 5e-5, it
        assert max_rel(st.grad_oT.cpu().numpy(), o64["grad_oT"].cpu().numpy()) < 1e-5, it
        # everything behind D4: the gradient into the target cloud, the summed parameter gradients (the bucket), the
        # parameters after SGD (through their accumulated update), BatchNorm running statistics
        now = st.d4.state_dict()
        items = [("grad_vertT", st.grad_vertT, o64["grad_vertT"], o32["grad_vertT"])]
        items += [("grad:" + k, v, o64["grads"][k], o32["grads"][k]) for k, v in zip(st._pnames, st.bucket.views)]
        items += [("update:" + k, now[k].double() - p_before[k], sd[torch.float64][k] - p_before[k],
                   sd[torch.float32][k].double() - p_before[k]) for k in st._pnames]
        worst = 0.0
        rows = []
        gmax = max(g.abs().max().item() for g in o64["grads"].values())
        dead = {k for k, g in o64["grads"].items() if g.abs().max().item() < 1e-6 * gmax}    # e.g. feat.bn3.bias: its shift is
        for name, a, b, b32 in items:                                                          # cancelled by the head's bn1
            k = name.split(":", 1)[-1]
            leaf = k.rsplit(".", 2)[-2] if "." in k else ""
            if k in dead or (k.endswith(".bias") and (leaf.startswith("conv") or leaf in ("fc1", "fc2"))):
                continue        # mathematically zero gradient (train-mode BN cancels the preceding bias): noise on all sides
            b = b.detach().double().cpu().numpy()
            if np.abs(b).max() < 1e-12:
                continue
            rows.append((name, _l2(a.detach().double().cpu().numpy(), b), _l2(b32.detach().double().cpu().numpy(), b)))
            worst = max(worst, rows[-1][2])
        print(f"step {shape} {precision} it={it}: worst oracle fp32-vs-fp64 distance {worst:.1e}; largest errors: " +
              ", ".join(f"{n} {e:.1e}/{f:.1e}" for n, e, f in sorted(rows, key=lambda r: -r[1])[:4]))
        for name, err, floor in rows:
            # update:* of step 2 contains fp32 parameter resolution: |p| 2^-24 against an update of lr |g|
            extra = 2e-3 if name.startswith("update:") else 0.0
            assert err <= max(base, 4 * floor, 2 * worst) + extra, (it, name, err, floor, worst)
        for k, v in now.items():
            if not _is_bn_buffer(k):
                continue
            if k.endswith(("running_mean", "running_var")):
                assert max_rel(v.cpu().numpy(), sd[torch.float64][k].cpu().numpy()) < (BF16_TOL if bf16 else 1e-4), (it, k)
            elif k.endswith("num_batches_tracked"):
                assert int(v.item()) == int(sd[torch.float64][k].item()) == 3 * (it + 1), (it, k, int(v.item()))
        # momentum buffer after the step
        off = 0
        for k, p_ in zip(st._pnames, st.bucket.params):
            mine = st.opt.momentum_buffer[off:off + p_.numel()].double().cpu().numpy()
            off += p_.numel()
            ref_m = mom[torch.float64][k].cpu().numpy().reshape(-1)
            f_m = _l2(mom[torch.float32][k].double().cpu().numpy(), ref_m)
            if np.abs(ref_m).max() > 1e-12:
                assert _l2(mine, ref_m) <= max(base, 4 * f_m, 2 * worst) + 1e-6, (it, "momentum:" + k)


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("shape", ["cfg2_small_maps", "mmwhs_shard"])
def test_shared_trunk_forward_is_bit_identical(shape, graph):
    """P4 taking the part of D4's forward pass that P2 already computed on the same cloud (share_forward: both trunks and
    the STN head, everything below the classifier head's Dropout) against every pass computing its own: same seeds, same
    Dropout draws, two steps — results, gradients, parameters, BatchNorm buffers and momentum must match bit for bit."""
    from pointcloududa_b200.step import AdversarialStep, StepConfig
    w = STEP_SHAPES[shape]
    cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], activation=w["activation"],
                     normalize=w["normalize"], return_prob=w["return_prob"], precision="bf16", lr_dis=2.5e-3)
    dev = torch.device(DEV)
    host = _step_inputs(w, seed=78)
    outs = []
    for share in (False, True):
        torch.cuda.manual_seed(1234)
        st = AdversarialStep(cfg, dev, seed=3, share_trunk=share)
        st.load_inputs(host, non_blocking=False)
        if graph:
            st.capture(warmup=2)
        torch.cuda.manual_seed(99)
        res = [st.run().clone() for _ in range(2)]
        torch.cuda.synchronize()
        outs.append((res, st.grad_vertT.clone(), st.bucket.flat.clone(), st.opt.momentum_buffer.clone(),
                     {k: v.clone() for k, v in st.d4.state_dict().items()}))
        st.close()
    (ra, ga, fa, ma, sa), (rb, gb, fb, mb, sb) = outs
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)
    assert torch.equal(ga, gb) and torch.equal(fa, fb) and torch.equal(ma, mb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k

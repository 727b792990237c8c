"""numpy restatement of the reference's point-cloud discriminator — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/networks/PointNetCls.py:
  STN3d.forward        :38-63    STNkd.forward :80-102
  PointNetfeat.forward :135-168  PointNetCls.forward :204-214
Train-mode BatchNorm1d throughout (the reference never calls .eval() on D4), biased variance for
normalisation, unbiased for the running buffers (torch.nn.BatchNorm1d, eps 1e-5, momentum 0.1).
Only the batch-size > 1 branches are restated: the B == 1 InstanceNorm branch of the reference
crashes on torch >= 2 (SURVEY.md §8c caveat 2).

The backward pass is a plain dense reverse-mode sweep written from the textbook formulas — it
deliberately shares no structure with the CUDA implementation's low-rank pooled-layer backward,
so agreement between the two is evidence, not tautology.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass(frozen=True)
class PointNetSpec:
    feature_transform: bool = False
    sample_transform: bool = True
    ext: bool = False
    in_channel: int = 3
    dim: int = 3


def _stn_shapes(prefix: str, k_in: int, k_out: int):
    s = {}
    for name, (co, ci) in {"conv1": (64, k_in), "conv2": (128, 64), "conv3": (1024, 128)}.items():
        s[f"{prefix}{name}.weight"] = (co, ci, 1)
        s[f"{prefix}{name}.bias"] = (co,)
    for name, (co, ci) in {"fc1": (512, 1024), "fc2": (256, 512), "fc3": (k_out, 256)}.items():
        s[f"{prefix}{name}.weight"] = (co, ci)
        s[f"{prefix}{name}.bias"] = (co,)
    for name, c in {"bn1": 64, "bn2": 128, "bn3": 1024, "bn4": 512, "bn5": 256}.items():
        s[f"{prefix}{name}.weight"] = (c,)
        s[f"{prefix}{name}.bias"] = (c,)
    return s


def param_shapes(spec: PointNetSpec) -> dict:
    """Trainable tensors of PointNetCls, keyed like its state_dict (PointNetCls.py:170-186)."""
    s = {}
    s.update(_stn_shapes("feat.stn.", spec.dim, 9))
    if spec.ext:
        convs = {"conv1": (8, spec.in_channel), "conv1_1": (64, 8), "conv2": (128, 64),
                 "conv2_1": (256, 128), "conv3": (512, 256), "conv3_1": (1024, 512)}
        bns = {"bn1": 8, "bn1_1": 64, "bn2": 128, "bn2_1": 256, "bn3": 512, "bn3_1": 1024}
    else:
        convs = {"conv1": (64, spec.in_channel), "conv2": (128, 64), "conv3": (1024, 128)}
        bns = {"bn1": 64, "bn2": 128, "bn3": 1024}
    for name, (co, ci) in convs.items():
        s[f"feat.{name}.weight"] = (co, ci, 1)
        s[f"feat.{name}.bias"] = (co,)
    for name, c in bns.items():
        s[f"feat.{name}.weight"] = (c,)
        s[f"feat.{name}.bias"] = (c,)
    if spec.feature_transform:
        s.update(_stn_shapes("feat.fstn.", 64, 64 * 64))
    for name, (co, ci) in {"fc1": (512, 1024), "fc2": (256, 512), "fc3": (1, 256)}.items():
        s[f"{name}.weight"] = (co, ci)
        s[f"{name}.bias"] = (co,)
    for name, c in {"bn1": 512, "bn2": 256}.items():
        s[f"{name}.weight"] = (c,)
        s[f"{name}.bias"] = (c,)
    return s


def make_params(spec: PointNetSpec, seed: int) -> dict:
    """Deterministic, torch-RNG-independent parameters (same law as torch's default init:
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear; BN gamma perturbed around 1 and beta around
    0 so that gamma's sign and beta actually matter in the tests)."""
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in sorted(param_shapes(spec).items()):
        mod = key.rsplit(".", 1)[0].rsplit(".", 1)[-1]
        if mod.startswith("bn"):
            if key.endswith("weight"):
                v = 1.0 + 0.5 * rs.standard_normal(shape)  # some gammas go negative / near 0
            else:
                v = 0.2 * rs.standard_normal(shape)
        else:
            fan_in = shape[1] if len(shape) > 1 else None
            if fan_in is None:  # bias: fan_in of the matching weight
                wshape = param_shapes(spec)[key[:-4] + "weight"]
                fan_in = wshape[1]
            bound = 1.0 / np.sqrt(fan_in)
            v = rs.uniform(-bound, bound, size=shape)
        out[key] = v.astype(np.float32)
    return out


# ---- primitive layers (forward returns cache; backward consumes it) ----------------------------
def _bn_fwd(y, gamma, beta, axes, dt):
    mean = y.mean(axis=axes, keepdims=True, dtype=dt)
    var = ((y - mean) ** 2).mean(axis=axes, keepdims=True, dtype=dt)
    invstd = 1.0 / np.sqrt(var + dt(BN_EPS))
    yhat = (y - mean) * invstd
    shape = [1] * y.ndim
    shape[1] = -1
    z = yhat * gamma.reshape(shape) + beta.reshape(shape)
    return z, (yhat, invstd, gamma.reshape(shape), axes, mean.reshape(-1), var.reshape(-1))


def _bn_bwd(dz, cache):
    yhat, invstd, g, axes, _, _ = cache
    dgamma = (dz * yhat).sum(axis=axes)
    dbeta = dz.sum(axis=axes)
    m = 1
    for a in axes:
        m *= dz.shape[a]
    dyhat = dz * g
    dy = invstd * (dyhat - dyhat.mean(axis=axes, keepdims=True)
                   - yhat * (dyhat * yhat).mean(axis=axes, keepdims=True))
    return dy, dgamma, dbeta


def bf16_round(a):
    """Round-to-nearest-even to bfloat16 precision (returned in a's dtype).  Restates the operand
    rounding of the product's tensor-core path (bf16 operands, fp32 accumulation) so that tests can
    separate that rounding -- which train-mode BatchNorm over a handful of samples amplifies -- from
    genuine kernel errors."""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a32.shape).astype(a.dtype)


def bf16_layer(cin: int) -> bool:
    """Layers the product runs on the tensor cores: contraction length a multiple of 64 in [64, 512]."""
    return 64 <= cin <= 512 and cin % 64 == 0


def _conv_bn(x2, W, b, gamma, beta, relu, dt, bf16_operands=False):
    """x2 [Cin, B*N] -> relu?(bn(W x + b)) [Cout, B*N]  (Conv1d k=1 + BatchNorm1d, :41-43,:143-161).
    The k=1 convolution over [B,C,N] is one GEMM over the B*N columns."""
    W2 = W.reshape(W.shape[0], -1).astype(dt)
    if bf16_operands and bf16_layer(W2.shape[1]):
        W2, x2 = bf16_round(W2), bf16_round(x2)
    y = W2 @ x2 + b.astype(dt)[:, None]
    mean = y.mean(axis=1, keepdims=True, dtype=dt)
    var = ((y - mean) ** 2).mean(axis=1, keepdims=True, dtype=dt)
    invstd = 1.0 / np.sqrt(var + dt(BN_EPS))
    yhat = (y - mean) * invstd
    g = gamma.astype(dt)[:, None]
    z = yhat * g + beta.astype(dt)[:, None]
    a = np.maximum(z, 0) if relu else z
    return a, (x2, W2, (yhat, invstd, g, (1,), mean.reshape(-1), var.reshape(-1)), relu, z)


def _conv_bn_bwd(da, cache):
    x2, W2, bnc, relu, z = cache
    yhat, invstd, g = bnc[0], bnc[1], bnc[2]
    dz = da * (z > 0) if relu else da
    dgamma = (dz * yhat).sum(axis=1)
    dbeta = dz.sum(axis=1)
    dyhat = dz * g
    dy = invstd * (dyhat - dyhat.mean(axis=1, keepdims=True) - yhat * (dyhat * yhat).mean(axis=1, keepdims=True))
    dW = dy @ x2.T
    db = dy.sum(axis=1)
    dx = W2.T @ dy
    return dx, {"weight": dW, "bias": db, "gamma": dgamma, "beta": dbeta}


def _fc(x, W, b, dt):
    return x @ W.astype(dt).T + b.astype(dt), (x, W.astype(dt))


def _fc_bwd(dy, cache):
    x, W = cache
    return dy @ W, {"weight": dy.T @ x, "bias": dy.sum(axis=0)}


def _fc_bn_relu(x, W, b, gamma, beta, dt, mask=None):
    y, fcc = _fc(x, W, b, dt)
    if mask is not None:  # Dropout sits between fc2 and bn2 in the head (:209)
        y = y * mask
    z, bnc = _bn_fwd(y, gamma.astype(dt), beta.astype(dt), (0,), dt)
    return np.maximum(z, 0), (fcc, bnc, z, mask)


def _fc_bn_relu_bwd(da, cache):
    fcc, bnc, z, mask = cache
    dz = da * (z > 0)
    dy, dgamma, dbeta = _bn_bwd(dz, bnc)
    if mask is not None:
        dy = dy * mask
    dx, g = _fc_bwd(dy, fcc)
    g.update({"gamma": dgamma, "beta": dbeta})
    return dx, g


# ---- shared-MLP stack + max-pool: the op the CUDA library implements ----------------------------
def stack_forward(x, layers, pool: bool, dtype=np.float32, bf16_operands=False):
    """layers: list of dict(weight [Cout,Cin(,1)], bias, gamma, beta, relu).  x [B,C0,N].
    Returns (out, cache); out is [B,C_L] when pool else [B,C_L,N].  cache['stats'] holds the
    per-layer batch mean / biased var (for running-stat checks)."""
    dt = dtype
    B, C0, N = x.shape
    a = np.ascontiguousarray(x.astype(dt).transpose(1, 0, 2)).reshape(C0, B * N)
    caches, stats = [], []
    for L in layers:
        a, c = _conv_bn(a, L["weight"], L["bias"], L["gamma"], L["beta"], bool(L["relu"]), dt, bf16_operands)
        caches.append(c)
        stats.append((c[2][4], c[2][5]))
    a3 = a.reshape(a.shape[0], B, N).transpose(1, 0, 2)  # [B, C, N]
    arg = None
    if pool:
        arg = a3.argmax(axis=2)  # first maximal index, like torch.max(x, 2)
        out = np.take_along_axis(a3, arg[:, :, None], axis=2)[:, :, 0]
    else:
        out = np.ascontiguousarray(a3)
    return out, {"caches": caches, "arg": arg, "shape": a3.shape, "stats": stats, "pool": pool}


def stack_backward(dout, cache):
    """Returns (dx [B,C0,N], [per-layer dict(weight,bias,gamma,beta)])."""
    B, C, N = cache["shape"]
    if cache["pool"]:
        da3 = np.zeros(cache["shape"], dout.dtype)
        np.put_along_axis(da3, cache["arg"][:, :, None], dout[:, :, None], axis=2)
    else:
        da3 = dout
    da = np.ascontiguousarray(da3.transpose(1, 0, 2)).reshape(C, B * N)
    grads = []
    for c in reversed(cache["caches"]):
        da, g = _conv_bn_bwd(da, c)
        grads.append(g)
    dx = np.ascontiguousarray(da.reshape(da.shape[0], B, N).transpose(1, 0, 2))
    return dx, grads[::-1]


def _stn_layers(params, prefix):
    return [dict(weight=params[f"{prefix}conv{i}.weight"], bias=params[f"{prefix}conv{i}.bias"],
                 gamma=params[f"{prefix}bn{i}.weight"], beta=params[f"{prefix}bn{i}.bias"], relu=True)
            for i in (1, 2, 3)]


def feat_stacks(params, spec: PointNetSpec):
    """The trunk of PointNetfeat as (stack before the feature transform, stack after it)."""
    def L(conv, bn, relu):
        return dict(weight=params[f"feat.{conv}.weight"], bias=params[f"feat.{conv}.bias"],
                    gamma=params[f"feat.{bn}.weight"], beta=params[f"feat.{bn}.bias"], relu=relu)
    if spec.ext:
        pre = [L("conv1", "bn1", True), L("conv1_1", "bn1_1", True)]
        post = [L("conv2", "bn2", True), L("conv2_1", "bn2_1", True), L("conv3", "bn3", False),
                L("conv3_1", "bn3_1", True)]
    else:
        pre = [L("conv1", "bn1", True)]
        post = [L("conv2", "bn2", True), L("conv3", "bn3", False)]
    return pre, post


def _stn_forward(params, prefix, x, k, dt, bf16_operands=False):
    g, sc = stack_forward(x, _stn_layers(params, prefix), True, dt, bf16_operands)
    h1, c1 = _fc_bn_relu(g, params[f"{prefix}fc1.weight"], params[f"{prefix}fc1.bias"],
                         params[f"{prefix}bn4.weight"], params[f"{prefix}bn4.bias"], dt)
    h2, c2 = _fc_bn_relu(h1, params[f"{prefix}fc2.weight"], params[f"{prefix}fc2.bias"],
                         params[f"{prefix}bn5.weight"], params[f"{prefix}bn5.bias"], dt)
    t, c3 = _fc(h2, params[f"{prefix}fc3.weight"], params[f"{prefix}fc3.bias"], dt)
    t = t + np.eye(k, dtype=dt).reshape(1, k * k)
    return t.reshape(-1, k, k), (sc, c1, c2, c3, prefix)


def _stn_backward(dtrans, cache, grads):
    sc, c1, c2, c3, prefix = cache
    B = dtrans.shape[0]
    d, g3 = _fc_bwd(dtrans.reshape(B, -1), c3)
    d, g2 = _fc_bn_relu_bwd(d, c2)
    d, g1 = _fc_bn_relu_bwd(d, c1)
    dx, gs = stack_backward(d, sc)
    for name, g in (("fc3", g3),):
        grads[f"{prefix}{name}.weight"] = g["weight"]; grads[f"{prefix}{name}.bias"] = g["bias"]
    for name, bn, g in (("fc1", "bn4", g1), ("fc2", "bn5", g2)):
        grads[f"{prefix}{name}.weight"] = g["weight"]; grads[f"{prefix}{name}.bias"] = g["bias"]
        grads[f"{prefix}{bn}.weight"] = g["gamma"]; grads[f"{prefix}{bn}.bias"] = g["beta"]
    for i, g in zip((1, 2, 3), gs):
        grads[f"{prefix}conv{i}.weight"] = g["weight"][:, :, None]
        grads[f"{prefix}conv{i}.bias"] = g["bias"]
        grads[f"{prefix}bn{i}.weight"] = g["gamma"]; grads[f"{prefix}bn{i}.bias"] = g["beta"]
    return dx


def pointnet_forward(params, x, spec: PointNetSpec = PointNetSpec(), drop_mask=None,
                     dtype=np.float32, bf16_operands=False):
    """x [B,3,N] -> (logit [B,1], trans [B,3,3]|None, trans_feat [B,64,64]|None, cache).
    drop_mask: the (already 1/(1-p)-scaled) dropout mask applied to fc2's output, or None.
    bf16_operands: round the operands of the wide shared-MLP layers to bf16 (see bf16_round)."""
    dt = dtype
    x = x.astype(dt)
    cache = {"spec": spec}
    trans = None
    if spec.sample_transform:
        trans, cache["stn"] = _stn_forward(params, "feat.stn.", x, 3, dt, bf16_operands)
        cache["x_in"] = x
        x = np.matmul(x.transpose(0, 2, 1), trans).transpose(0, 2, 1)  # :140-142
        cache["trans"] = trans
    pre, post = feat_stacks(params, spec)
    h, cache["pre"] = stack_forward(x, pre, False, dt, bf16_operands)
    trans_feat = None
    if spec.feature_transform:
        trans_feat, cache["fstn"] = _stn_forward(params, "feat.fstn.", h, 64, dt, bf16_operands)
        cache["h_in"] = h
        h = np.matmul(h.transpose(0, 2, 1), trans_feat).transpose(0, 2, 1)  # :147-151
        cache["trans_feat"] = trans_feat
    g, cache["post"] = stack_forward(h, post, True, dt, bf16_operands)
    a1, cache["h1"] = _fc_bn_relu(g, params["fc1.weight"], params["fc1.bias"], params["bn1.weight"],
                                  params["bn1.bias"], dt)
    a2, cache["h2"] = _fc_bn_relu(a1, params["fc2.weight"], params["fc2.bias"], params["bn2.weight"],
                                  params["bn2.bias"], dt, mask=drop_mask)
    logit, cache["h3"] = _fc(a2, params["fc3.weight"], params["fc3.bias"], dt)
    cache["global_feat"] = g
    return logit, trans, trans_feat, cache


def pointnet_backward(dlogit, cache):
    """Returns (dx [B,3,N], grads keyed like the state_dict)."""
    spec = cache["spec"]
    grads = {}
    d, g = _fc_bwd(dlogit, cache["h3"])
    grads["fc3.weight"], grads["fc3.bias"] = g["weight"], g["bias"]
    d, g = _fc_bn_relu_bwd(d, cache["h2"])
    grads["fc2.weight"], grads["fc2.bias"], grads["bn2.weight"], grads["bn2.bias"] = g["weight"], g["bias"], g["gamma"], g["beta"]
    d, g = _fc_bn_relu_bwd(d, cache["h1"])
    grads["fc1.weight"], grads["fc1.bias"], grads["bn1.weight"], grads["bn1.bias"] = g["weight"], g["bias"], g["gamma"], g["beta"]

    names_pre = [("conv1", "bn1"), ("conv1_1", "bn1_1")] if spec.ext else [("conv1", "bn1")]
    names_post = ([("conv2", "bn2"), ("conv2_1", "bn2_1"), ("conv3", "bn3"), ("conv3_1", "bn3_1")]
                  if spec.ext else [("conv2", "bn2"), ("conv3", "bn3")])

    def put(names, gs):
        for (conv, bn), g in zip(names, gs):
            grads[f"feat.{conv}.weight"] = g["weight"][:, :, None]
            grads[f"feat.{conv}.bias"] = g["bias"]
            grads[f"feat.{bn}.weight"] = g["gamma"]; grads[f"feat.{bn}.bias"] = g["beta"]

    dh, gs = stack_backward(d, cache["post"])
    put(names_post, gs)
    if spec.feature_transform:
        h_in, tf = cache["h_in"], cache["trans_feat"]
        # h' = (h^T tf)^T = tf^T h
        dtf = np.matmul(h_in, dh.transpose(0, 2, 1))          # [B,64,64]
        dh_direct = np.matmul(tf, dh)                          # [B,64,N]
        dh = dh_direct + _stn_backward(dtf, cache["fstn"], grads)
    dx, gs = stack_backward(dh, cache["pre"])
    put(names_pre, gs)
    if spec.sample_transform:
        x_in, tr = cache["x_in"], cache["trans"]
        dtr = np.matmul(x_in, dx.transpose(0, 2, 1))           # [B,3,3]
        dx = np.matmul(tr, dx) + _stn_backward(dtr, cache["stn"], grads)
    return dx, grads

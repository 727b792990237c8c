"""Plain-torch restatement of ONE adversarial train step of the reference — TEST INFRASTRUCTURE ONLY.

Restates, with stock torch ops on whatever device / dtype the caller chooses (float64 on the GPU is the
usual choice), the non-backbone work of `train_epoch` (/root/reference/src/train_mscmrseg.py:183-330;
train_mmwhs.py:187-360) with the backbone outputs given as inputs:

  P1 :206,:213   loss_seg3 = batch_NN_loss(vertS, vertexA), gradient into vertS
  P2 :222        uncertainty_mapT = -1.0 * sigmoid(oT) * log(sigmoid(oT) + smooth)       (MS-CMRSeg form)
     train_mmwhs.py:213-217,224-225,240-243   softmax form: p = softmax(oT, 1); -1.0 * p * log(p + smooth) / log(C)
     :230-233    loss_vert_target; D_out4 = model_dis4(vertT.transpose(2, 1))[0] with FROZEN parameters;
                 loss_adv_diff_point = dr * BCE_with_logits(D_out4, 1); backward into vertT (:246)
  P3 :284-291    D4(vertS.detach()) -> BCE vs 1 -> parameter gradients, accuracy
  P4 :314-322    D4(vertT.detach()) -> BCE vs 0 -> parameter gradients (accumulated), accuracy
  P5 :329-330    optim_dis4.step()  (SGD momentum .99, weight decay 5e-4, :450-455)

The discriminator is a functional restatement of networks/PointNetCls.py (STN3d :38-63, STNkd :80-102,
PointNetfeat :135-168, PointNetCls :204-214) over a state_dict, batch-size > 1 branches only; it is pinned
against the reference's own outputs by tests/test_oracle_golden.py (CPU).  `bf16_operands=True` rounds both
operands of every 1x1-conv layer the product runs on tensor cores (contraction length 64..512) to bfloat16
with a straight-through gradient, everything else stays in the working dtype: the same-rounding oracle for
the tcgen05 path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

RESULT_FIELDS = ("chamfer_source", "chamfer_target", "entropy_mean_target", "adv_point_loss",
                 "d4_loss_source", "d4_loss_target", "d4_acc_source", "d4_acc_target")


def _bf16_st(t: torch.Tensor) -> torch.Tensor:
    """Round to bfloat16 (nearest even), gradient passes straight through."""
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _tc_layer(cin: int) -> bool:
    return 64 <= cin <= 512 and cin % 64 == 0


class D4:
    """Functional PointNetCls over a dict of tensors keyed like the reference's state_dict."""

    def __init__(self, sd: Dict[str, torch.Tensor], feature_transform=False, ext=False, sample_transform=True,
                 bf16_operands=False, drop_masks=None):
        self.sd = sd
        self.ft, self.ext, self.st = feature_transform, ext, sample_transform
        self.bf16 = bf16_operands
        self.drop_masks = drop_masks            # iterator of [B,256] masks (already scaled by 1/(1-p)) or None
        self.pool_margin = []                   # per pooled stack: (top1 - top2) [B, C] of the pooled tensor

    # -- layers --------------------------------------------------------------------------------
    def _bn(self, y, name):
        sd = self.sd
        return F.batch_norm(y, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                            sd[name + ".bias"], training=True, momentum=0.1, eps=1e-5)

    def _bump(self, name):
        k = name + ".num_batches_tracked"
        if k in self.sd:
            self.sd[k] += 1

    def _conv(self, x, conv, bn, relu=True):
        w, b = self.sd[conv + ".weight"], self.sd[conv + ".bias"]
        if self.bf16 and _tc_layer(w.shape[1]):
            w, x = _bf16_st(w), _bf16_st(x)
        y = self._bn(F.conv1d(x, w, b), bn)
        self._bump(bn)
        return F.relu(y) if relu else y

    def _fc(self, x, fc, bn=None, relu=False, mask=None):
        y = F.linear(x, self.sd[fc + ".weight"], self.sd[fc + ".bias"])
        if mask is not None:
            y = y * mask
        if bn is not None:
            y = self._bn(y, bn)
            self._bump(bn)
        return F.relu(y) if relu else y

    def _pool(self, x):
        with torch.no_grad():
            top2 = x.topk(min(2, x.shape[2]), dim=2)[0]
            self.pool_margin.append(top2[..., 0] - top2[..., -1])
        return x.max(dim=2)[0]

    def _stn(self, x, p, k):
        x = self._conv(x, p + "conv1", p + "bn1")
        x = self._conv(x, p + "conv2", p + "bn2")
        x = self._conv(x, p + "conv3", p + "bn3")
        x = self._pool(x)
        x = self._fc(x, p + "fc1", p + "bn4", relu=True)
        x = self._fc(x, p + "fc2", p + "bn5", relu=True)
        x = self._fc(x, p + "fc3")
        x = x + torch.eye(k, dtype=x.dtype, device=x.device).reshape(1, k * k)
        return x.view(-1, k, k)

    def __call__(self, x):
        """x [B, 3, N] -> (logit [B,1], trans, trans_feat)"""
        trans = trans_feat = None
        if self.st:
            trans = self._stn(x, "feat.stn.", 3)
            x = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)
        x = self._conv(x, "feat.conv1", "feat.bn1")
        if self.ext:
            x = self._conv(x, "feat.conv1_1", "feat.bn1_1")
        if self.ft:
            trans_feat = self._stn(x, "feat.fstn.", 64)
            x = torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1)
        x = self._conv(x, "feat.conv2", "feat.bn2")
        if self.ext:
            x = self._conv(x, "feat.conv2_1", "feat.bn2_1")
        x = self._conv(x, "feat.conv3", "feat.bn3", relu=False)
        if self.ext:
            x = self._conv(x, "feat.conv3_1", "feat.bn3_1")
        x = self._pool(x)
        x = self._fc(x, "fc1", "bn1", relu=True)
        mask = next(self.drop_masks) if self.drop_masks is not None else None
        x = self._fc(x, "fc2", "bn2", relu=True, mask=mask)
        return self._fc(x, "fc3"), trans, trans_feat


def nn_loss(x: torch.Tensor, y: torch.Tensor):
    """utils/loss.py:40-76 (N == M): Euclidean nearest-neighbour distances in both directions."""
    rx = (x * x).sum(-1)
    ry = (y * y).sum(-1)
    P = rx[:, :, None] + ry[:, None, :] - 2.0 * torch.bmm(x, y.transpose(2, 1))
    d = torch.sqrt(P + 0.00001)
    v1, i1 = d.min(dim=2)
    v2, i2 = d.min(dim=1)
    B, N = x.shape[0], x.shape[1]
    return v1.sum(1).div(N).sum().div(B) + v2.sum(1).div(N).sum().div(B), i1, i2


def entropy_map(z: torch.Tensor, activation: str, normalize: bool, smooth: float = 1e-7):
    p = torch.sigmoid(z) if activation == "sigmoid" else torch.softmax(z, dim=1)
    m = -1.0 * p * torch.log(p + smooth)
    if normalize:
        m = m / math.log(z.shape[1])
    return m, p


def trainable_keys(sd):
    """Parameters of PointNetCls in state_dict order (the in1..in5 InstanceNorm modules are affine-free)."""
    return [k for k in sd if k.endswith((".weight", ".bias"))]


def adversarial_step(sd: Dict[str, torch.Tensor], inputs: Dict[str, torch.Tensor], momentum_buf: Dict[str, torch.Tensor],
                     activation="sigmoid", normalize=False, dr=0.01, lr=2.5e-5, momentum=0.99, weight_decay=0.0005,
                     feature_transform=False, ext=False, bf16_operands=False, world=1,
                     param_keys: Optional[list] = None):
    """Runs one step IN PLACE on `sd` (parameters + BatchNorm buffers) and `momentum_buf`.
    Returns dict(result=[8], grad_vertS, grad_vertT, grad_oT, grads={param: dL/dp (sum of P3 and P4) / world})."""
    keys = param_keys if param_keys is not None else trainable_keys(sd)
    kw = dict(feature_transform=feature_transform, ext=ext, bf16_operands=bf16_operands)
    out = {}
    # P1
    vS = inputs["vertS"].detach().requires_grad_(True)
    l3, _, _ = nn_loss(vS, inputs["vertexA"])
    (out["grad_vertS"],) = torch.autograd.grad(l3, vS)
    # P2
    oT = inputs["oT"].detach().requires_grad_(True)
    mT, _ = entropy_map(oT, activation, normalize)
    (out["grad_oT"],) = torch.autograd.grad(mT, oT, inputs["grad_mapT"])
    ent_mean = mT.detach().sum(dim=1).mean()
    with torch.no_grad():
        lvt, _, _ = nn_loss(inputs["vertT"], inputs["vertexB"])
    vT = inputs["vertT"].detach().requires_grad_(True)
    frozen = {k: v.detach() for k, v in sd.items()}
    margins = []
    net = D4(frozen, **kw)
    d_out = net(vT.transpose(2, 1))[0]
    margins.append(net.pool_margin)
    l_adv = dr * F.binary_cross_entropy_with_logits(d_out, torch.ones_like(d_out))
    (out["grad_vertT"],) = torch.autograd.grad(l_adv, vT)
    # P3 / P4
    leaves = {k: (v.detach().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    net = D4(leaves, **kw)
    d_s = net(inputs["vertS"].detach().transpose(2, 1))[0]
    margins.append(net.pool_margin)
    l_s = F.binary_cross_entropy_with_logits(d_s, torch.ones_like(d_s))
    net = D4(leaves, **kw)
    d_t = net(inputs["vertT"].detach().transpose(2, 1))[0]
    margins.append(net.pool_margin)
    l_t = F.binary_cross_entropy_with_logits(d_t, torch.zeros_like(d_t))
    gs = torch.autograd.grad(l_s + l_t, [leaves[k] for k in keys], allow_unused=True)
    acc_s = (torch.sigmoid(d_s.detach()) >= 0.5).to(d_s.dtype).mean()
    acc_t = 1.0 - (torch.sigmoid(d_t.detach()) >= 0.5).to(d_t.dtype).mean()
    grads = {k: (torch.zeros_like(sd[k]) if g is None else g / world) for k, g in zip(keys, gs)}
    # P5: torch.optim.SGD arithmetic (dampening 0, nesterov off)
    first = not momentum_buf
    with torch.no_grad():
        for k in keys:
            g = grads[k] + weight_decay * sd[k]
            if first or k not in momentum_buf:
                momentum_buf[k] = g.clone()
            else:
                momentum_buf[k].mul_(momentum).add_(g)
            sd[k].sub_(lr * momentum_buf[k])
    out["grads"] = grads
    out["pool_margins"] = margins
    out["logits"] = (d_out.detach(), d_s.detach(), d_t.detach())
    out["result"] = torch.stack([l3.detach(), lvt, ent_mean, l_adv.detach(), l_s.detach(), l_t.detach(), acc_s, acc_t])
    return out


def conditioned_inputs(w: dict, seed: int) -> Dict[str, torch.Tensor]:
    """Synthetic step inputs (SURVEY.md §8d shapes) whose clouds differ in extent per sample.  With i.i.d.
    unit-cube clouds the pooled features of all samples nearly coincide and the batch-of-B BatchNorms of the
    FC heads amplify rounding by orders of magnitude (see oracle/gen_golden.py): that would measure the
    network's conditioning, not the implementation under test."""
    from pointcloududa_b200.synth import synth_clouds, synth_logits
    B, C, H, W, N = w["B"], w["C"], w["H"], w["W"], w["N"]
    xS, yA = synth_clouds(B, N, seed=seed + 1, mode="near")
    xT, yB = synth_clouds(B, N, seed=seed + 2, mode="uniform")
    g = torch.Generator().manual_seed(seed)
    sc = torch.rand(B, 1, 3, generator=g) * 0.7 + 0.3
    return {"oS": synth_logits(B, C, H, W, seed=seed + 3), "oT": synth_logits(B, C, H, W, seed=seed + 4),
            "grad_mapT": synth_logits(B, C, H, W, seed=seed + 5, std=1.0),
            "vertS": xS * sc, "vertT": xT * sc.flip(0), "vertexA": yA * sc, "vertexB": yB * sc.flip(0)}

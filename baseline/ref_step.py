"""One adversarial step executed by the REFERENCE'S OWN CODE under torch eager — the baseline arm of bench.py.

Runs the stock code path of sulaimanvesal/PointCloudUDA for the hot path, from the unmodified files installed in
`baseline/_ref/` (baseline/make_ref.py):

  * `utils.loss.batch_NN_loss`            (baseline/_ref/src/utils/loss.py:40-76)
  * `networks.PointNetCls.PointNetCls`    (baseline/_ref/src/networks/PointNetCls.py:170-214)
  * the inline entropy-map expressions, restated literally from `train_mscmrseg.py:222,265` and
    `train_mmwhs.py:213-217,224-225,240-243` (they are expressions inside the training loop, not functions)

in the order of `train_epoch` (train_mscmrseg.py:183-330) with the backbone outputs given as inputs, D4's SGD
(momentum .99, weight decay 5e-4, :450-455) included.  None of this repository's kernels, models or engine is on
this path.  On CUDA tensors the reference runs unchanged; on the CPU the one token `torch.cuda.LongTensor`
(loss.py:59) is replaced IN MEMORY by `torch.LongTensor` — as written the function cannot run without a GPU
(SURVEY.md §8c caveat 1).
"""
from __future__ import annotations

import math
import sys
import types
from pathlib import Path

REF = Path(__file__).resolve().parent / "_ref" / "src"
SMOOTH = 1e-7            # `smooth` of train_mscmrseg.py / train_mmwhs.py


def available() -> bool:
    return (REF / "utils" / "loss.py").exists() and (REF / "networks" / "PointNetCls.py").exists()


def load_reference(device_type: str):
    """(batch_NN_loss, PointNetCls) of the reference."""
    src = (REF / "utils" / "loss.py").read_text()
    if device_type != "cuda":
        if src.count("torch.cuda.LongTensor") != 1:
            raise RuntimeError("unexpected reference loss.py")
        src = src.replace("torch.cuda.LongTensor", "torch.LongTensor")
    mod = types.ModuleType("ref_utils_loss")
    exec(compile(src, str(REF / "utils" / "loss.py"), "exec"), mod.__dict__)
    sys.path.insert(0, str(REF))
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "networks" or k.startswith("networks.")}
        from networks.PointNetCls import PointNetCls  # type: ignore
        for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
            sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    finally:
        sys.path.pop(0)
    return mod.batch_NN_loss, PointNetCls


class ReferenceStep:
    """Holds the reference's D4 + its optimiser and runs the hot-path part of one train_epoch iteration."""

    def __init__(self, w: dict, device, seed: int = 0, lr_dis: float = 2.5e-5, dr: float = 0.01):
        import torch
        self.torch = torch
        self.w = w
        self.device = torch.device(device)
        self.batch_NN_loss, PointNetCls = load_reference(self.device.type)
        torch.manual_seed(seed)
        self.model_dis4 = PointNetCls(feature_transform=bool(w.get("feature_transform", False)),
                                      ext=bool(w.get("ext", False))).to(self.device).train()
        self.optim_dis4 = torch.optim.SGD(self.model_dis4.parameters(), lr=lr_dis, momentum=0.99, weight_decay=0.0005)
        self.dr = dr

    def _map(self, o):
        torch = self.torch
        if self.w["activation"] == "sigmoid":                       # train_mscmrseg.py:222 / :265
            m = -1.0 * torch.sigmoid(o) * torch.log(torch.sigmoid(o) + SMOOTH)
            return m, None
        pred = torch.softmax(o, dim=1)                              # train_mmwhs.py:213 / :240
        m = -1.0 * pred * torch.log(pred + SMOOTH)
        if self.w["normalize"]:
            m = m / math.log(o.shape[1])                            # :224 / :242
        return m, pred

    def run(self, I: dict):
        """I: oS, oT, grad_mapT [B,C,H,W]; vertS, vertT, vertexA, vertexB [B,N,3] on self.device."""
        torch = self.torch
        import torch.nn.functional as F
        d4, dev = self.model_dis4, self.device
        self.optim_dis4.zero_grad()
        for p in d4.parameters():
            p.requires_grad = False
        # 1. (:206,:213) point loss of the source prediction, backward into the predicted cloud
        vertS = I["vertS"].detach().requires_grad_(True)
        loss_seg3 = self.batch_NN_loss(x=vertS, y=I["vertexA"])
        loss_seg3.backward()
        # 2. (:222-246) fool the discriminators: entropy map of the target logits (backward = what D2 sends back),
        #    target point loss (logged), D4 with frozen parameters
        oT = I["oT"].detach().requires_grad_(True)
        uncertainty_mapT, _ = self._map(oT)
        ent_mean = torch.mean(torch.sum(uncertainty_mapT.detach(), dim=1))     # train_mmwhs.py:243
        vertT = I["vertT"].detach().requires_grad_(True)
        loss_vert_target = self.batch_NN_loss(x=vertT, y=I["vertexB"])
        D_out4 = d4(vertT.transpose(2, 1))[0]
        loss_adv_diff_point = self.dr * F.binary_cross_entropy_with_logits(
            D_out4, torch.full(D_out4.shape, 1.0, device=dev))
        uncertainty_mapT.backward(I["grad_mapT"])
        loss_adv_diff_point.backward()
        # 3. (:250-291) discriminators on the source domain
        for p in d4.parameters():
            p.requires_grad = True
        with torch.no_grad():
            self._map(I["oS"])                                                   # :265 (input of D2, out of scope)
        D_out4 = d4(I["vertS"].detach().transpose(2, 1))[0]
        loss_D_same4 = F.binary_cross_entropy_with_logits(D_out4, torch.full(D_out4.shape, 1.0, device=dev))
        loss_D_same4.backward()
        acc1 = (torch.sigmoid(D_out4.detach()) >= .5).float().mean()
        # 4. (:314-322) discriminators on the target domain
        D_out4 = d4(I["vertT"].detach().transpose(2, 1))[0]
        loss_D_diff_4 = F.binary_cross_entropy_with_logits(D_out4, torch.full(D_out4.shape, 0.0, device=dev))
        loss_D_diff_4.backward()
        acc2 = 1 - (torch.sigmoid(D_out4.detach()) >= .5).float().mean()
        # 5. (:329-330)
        self.optim_dis4.step()
        self.grad_vertS, self.grad_vertT, self.grad_oT = vertS.grad, vertT.grad, oT.grad
        return torch.stack([loss_seg3.detach(), loss_vert_target.detach(), ent_mean, loss_adv_diff_point.detach(),
                            loss_D_same4.detach(), loss_D_diff_4.detach(), acc1, acc2])


def chamfer_forward(batch_NN_loss, x, y):
    return batch_NN_loss(x=x, y=y)

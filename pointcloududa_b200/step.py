"""The hot-path work of ONE adversarial train step, with the backbone outputs given as inputs.

Mirrors the non-backbone part of `train_epoch` (train_mscmrseg.py:183-330, train_mmwhs.py:187-360;
SURVEY.md §3.1 / §8d "one adversarial step"):

  P1  Chamfer(vertS, vertexA) forward + backward into the predicted source cloud        (:206,:213)
  P2  entropy map of the target logits, forward + backward (upstream gradient = what D2 would
      send back, supplied as an input because the D2/D1 conv stacks are out of scope)     (:222,:246)
      Chamfer(vertT, vertexB) forward only (logged)                                       (:230-231)
      D4(vertT^T) with FROZEN parameters -> BCE vs 1 -> backward into the cloud           (:232-233,:246)
  P3  entropy map of the detached source logits (forward only)                            (:265)
      D4(vertS.detach()^T) -> BCE vs 1 -> backward (parameter gradients)                  (:284-288)
  P4  D4(vertT.detach()^T) -> BCE vs 0 -> backward (parameter gradients)                  (:314-319)
  P5  [multi-GPU: all-reduce of D4's gradients]  SGD(momentum .99, wd 5e-4) step          (:329-330,:450-455)

The U-Net, D1/D2 and host-side bookkeeping are excluded.  All scalar results are kept on the
device and returned as one small tensor so a caller pays a single device->host read per step.

The three D4 passes (P2, P3, P4) do not depend on each other (D4's parameters only change in P5), nor do
the discriminator-free ops (Chamfer, entropy map), and at the reference's shapes (2400 points per D4 call)
no single kernel fills 148 SMs, so the step issues them on four streams: inside a captured CUDA graph they
become four concurrent branches (one per D4 pass + one for Chamfer / entropy).  The only shared
mutable state, BatchNorm's running statistics / num_batches_tracked, is recorded per pass and applied
after the join in the reference's pass order (networks.PointNetCls.deferred_bn_updates), so buffers and
gradients are identical to sequential execution.
The whole step is CUDA-graph capturable (`AdversarialStep.capture()`): no host synchronisation,
no allocation outside torch's graph pool, every libpcuda launch on the capture stream.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import dist as pdist
from .networks.PointNetCls import ForwardMemo, PointNetCls, apply_bn_updates, deferred_bn_updates, share_forward
from .ops import bce_with_logits_grad
from .optim import FusedSGD
from .utils.loss import batch_NN_loss, entropy_map

RESULT_FIELDS = ("chamfer_source", "chamfer_target", "entropy_mean_target", "adv_point_loss",
                 "d4_loss_source", "d4_loss_target", "d4_acc_source", "d4_acc_target")


@dataclass
class StepConfig:
    B: int
    C: int
    H: int
    W: int
    N: int
    activation: str = "sigmoid"   # MS-CMRSeg: sigmoid, no /log C;  MM-WHS -softmax: softmax, /log C
    normalize: bool = False
    return_prob: bool = False     # MM-WHS: D1 consumes pred (train_mmwhs.py:267)
    dr: float = 0.01              # args.dr weight of the adversarial terms
    lr_dis: float = 2.5e-5
    precision: str = "bf16"
    feature_transform: bool = False
    ext: bool = False
    sync_bn: bool = False         # multi-GPU: BatchNorm statistics over the clouds of all ranks instead of per rank


class AdversarialStep:
    """Owns D4, its optimiser, static input buffers and (optionally) a captured CUDA graph."""

    def __init__(self, cfg: StepConfig, device: torch.device, seed: int = 0, concurrent: bool = True,
                 exchange: str = "auto", share_trunk: bool = True):
        """exchange: how D4's gradient bucket is summed over the ranks of an initialised process group --
        "p2p": libpcuda's NVLink peer-memory kernel (pcuda_comm_allreduce_p2p; the bucket is packed straight into the
        communicator's symmetric buffer), "nccl": ncclAllReduce through libpcuda's own communicator, "torch":
        torch.distributed.all_reduce between two captured half-graphs (round-1 behaviour), "auto": p2p when peer memory
        is available, else nccl.  p2p / nccl are captured inside the ONE graph of the step.  "local": ignore the process
        group (a single-process step inside a multi-rank job, e.g. as the reference of a parity test).
        share_trunk: P2 and P4 see the same cloud with the same weights, so P4 takes the part of D4's forward pass below
        the classifier head's Dropout from P2 instead of recomputing it (bit-identical results either way)."""
        self.cfg = cfg
        self.device = device
        torch.manual_seed(seed)
        self.d4 = PointNetCls(feature_transform=cfg.feature_transform, ext=cfg.ext,
                              precision=cfg.precision).to(device).train()
        # optim_dis4 of the reference: SGD, momentum .99, weight decay 5e-4 (train_mscmrseg.py:450-455), as one
        # libpcuda launch over all parameters (optim.FusedSGD: torch's arithmetic, momentum buffers in one flat tensor)
        self.opt = FusedSGD(list(self.d4.parameters()), lr=cfg.lr_dis, momentum=0.99, weight_decay=0.0005)
        self._pnames = [n for n, _ in self.d4.named_parameters()]
        self._world = 1
        self.comm: Optional[pdist.PcudaComm] = None
        self.exchange = "none"
        self._force_local = exchange == "local"
        self._sync_world()
        if self._world > 1 and device.type == "cuda":
            if exchange not in ("auto", "p2p", "nccl", "torch", "local"):
                raise ValueError(f"unknown exchange {exchange!r}")
            if exchange != "torch":
                numel = sum(p.numel() for p in self.d4.parameters())
                self.comm = pdist.PcudaComm(device, p2p_floats=numel if exchange in ("auto", "p2p") else 0)
                if exchange == "p2p" and not self.comm.p2p:
                    raise RuntimeError("exchange='p2p' requested but peer memory is not available between the ranks")
                self.exchange = "p2p" if self.comm.p2p else "nccl"
            else:
                self.exchange = "torch"
        if cfg.sync_bn and self._world > 1:
            if self.comm is None:
                raise ValueError("sync_bn needs libpcuda's communicator (exchange 'auto', 'p2p' or 'nccl')")
            self.d4.set_sync_bn(self.comm)
        if self.exchange == "p2p":
            numel = sum(p.numel() for p in self.d4.parameters())
            self.bucket = pdist.GradBucket(self.d4.parameters(), flat=self.comm.buf_in[:numel])
        else:
            self.bucket = pdist.GradBucket(self.d4.parameters())
        self._grad_final = self.bucket.flat          # where the summed gradient is read by SGD
        B, C, H, W, N = cfg.B, cfg.C, cfg.H, cfg.W, cfg.N
        f32 = dict(dtype=torch.float32, device=device)
        # all inputs of a step are views of ONE flat buffer (every view 256-byte aligned), so a caller can move a
        # whole step's inputs with a single copy (`inputs_flat`; `input_layout()` gives the same views of any
        # other flat tensor, e.g. a pinned host staging buffer)
        self._input_shapes = {"oS": (B, C, H, W), "oT": (B, C, H, W), "grad_mapT": (B, C, H, W),
                              "vertS": (B, N, 3), "vertT": (B, N, 3), "vertexA": (B, N, 3), "vertexB": (B, N, 3)}
        self.inputs_flat = torch.zeros(self.input_numel(), **f32)
        self.inputs: Dict[str, torch.Tensor] = self.input_layout(self.inputs_flat)
        self.result = torch.zeros(len(RESULT_FIELDS), **f32)
        self.grad_oT = torch.zeros(B, C, H, W, **f32)
        self.grad_vertS = torch.zeros(B, N, 3, **f32)
        self.grad_vertT = torch.zeros(B, N, 3, **f32)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_post: Optional[torch.cuda.CUDAGraph] = None
        self._res = []
        self._pending = None
        self.share_trunk = share_trunk
        self._side = None
        self.concurrent = concurrent

    # -- inputs -----------------------------------------------------------------------------------
    def _input_offsets(self):
        offs, off = {}, 0
        for k, shp in self._input_shapes.items():
            n = 1
            for d in shp:
                n *= d
            offs[k] = (off, n, shp)
            off += (n + 63) // 64 * 64          # 256-byte aligned views
        return offs, off

    def input_numel(self) -> int:
        return self._input_offsets()[1]

    def input_layout(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The step's input tensors as views of `flat` (float32, input_numel() elements)."""
        offs, total = self._input_offsets()
        if flat.numel() != total or flat.dtype != torch.float32:
            raise ValueError(f"expected a float32 tensor of {total} elements")
        return {k: flat[o:o + n].view(shp) for k, (o, n, shp) in offs.items()}

    def load_inputs(self, host: Dict[str, torch.Tensor], non_blocking: bool = True) -> int:
        """Host (pinned) -> device copy of one step's inputs. Returns bytes copied."""
        n = 0
        for k, dst in self.inputs.items():
            dst.copy_(host[k], non_blocking=non_blocking)
            n += dst.numel() * dst.element_size()
        return n

    def _set_requires_grad(self, flag: bool) -> None:
        for p in self.d4.parameters():
            p.requires_grad = flag

    def _d4_fresh(self, x: torch.Tensor):
        leaves = [p.detach().requires_grad_(True) for p in self.bucket.params]
        out = torch.func.functional_call(self.d4, dict(zip(self._pnames, leaves)), (x,))[0]
        return out, leaves

    # -- the step -----------------------------------------------------------------------------------
    def _pre(self) -> None:
        cfg, I = self.cfg, self.inputs
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = tuple(torch.cuda.Stream(device=dev) for _ in range(3))
        s1, s2, s3 = self._side if self.concurrent else (main, main, main)
        sink_p2, sink_p3, sink_p4 = [], [], []

        # (D4's gradients live in the flat bucket; it is overwritten at the join)
        if self.concurrent:
            for s in (s1, s2, s3):
                s.wait_stream(main)

        # ---- branch D (stream s3) — the ops with no discriminator in them: P1 source Chamfer forward + backward
        # into the predicted cloud, the target entropy map forward + backward (the part of P2 that feeds D2),
        # the target Chamfer value.  None of them depends on D4, so they leave the critical branch.
        self._set_requires_grad(False)
        with torch.cuda.stream(s3):
            vertS = I["vertS"].detach().requires_grad_(True)
            loss_seg3 = batch_NN_loss(x=vertS, y=I["vertexA"])
            (gS,) = torch.autograd.grad(loss_seg3, vertS)
            self.grad_vertS.copy_(gS)
            oT = I["oT"].detach().requires_grad_(True)
            outs = entropy_map(oT, activation=cfg.activation, normalize=cfg.normalize,
                               return_prob=cfg.return_prob, return_mean=True)
            mapT, ent_mean = outs[0], outs[-1]
            (gO,) = torch.autograd.grad(mapT, oT, I["grad_mapT"])
            self.grad_oT.copy_(gO)
            with torch.no_grad():
                loss_vert_target = batch_NN_loss(x=I["vertT"], y=I["vertexB"])

        # ---- branch A (this stream) — P2: fool the point-cloud discriminator (D4 frozen: gradient into the cloud
        # only).  Issued before P3 / P4 so that Dropout consumes the generator in the reference's order (P2, P3, P4);
        # the streams run concurrently regardless of the issue order.
        # P2 and P4 run D4 on the same cloud with the same weights: everything below the classifier head's Dropout is
        # computed once (P2 records, P4 takes aliases: networks/PointNetCls.py, share_forward)
        memo = ForwardMemo() if self.share_trunk else None
        vertT = I["vertT"].detach().requires_grad_(True)
        with deferred_bn_updates(sink_p2), share_forward(memo):
            D_out4 = self.d4(vertT.transpose(2, 1))[0]
        loss_adv_point, _, g_logit = bce_with_logits_grad(D_out4, 1.0, cfg.dr)
        (gT,) = torch.autograd.grad(D_out4, vertT, g_logit)
        self.grad_vertT.copy_(gT)
        self._set_requires_grad(True)

        # ---- branch B (stream s1) — P3: discriminators on the source domain (label 1).  D4 runs on fresh
        # leaf views of its parameters (functional_call): their gradient-accumulator nodes are created on
        # the stream the branch runs on, so the step is legal under CUDA-graph capture.
        self._set_requires_grad(True)
        with torch.cuda.stream(s1):
            with torch.no_grad():
                entropy_map(I["oS"], activation=cfg.activation, normalize=cfg.normalize, return_prob=cfg.return_prob)
            with deferred_bn_updates(sink_p3):
                D_out4, leaves = self._d4_fresh(I["vertS"].transpose(2, 1))
            loss_D_same4, acc_s, g_logit = bce_with_logits_grad(D_out4, 1.0)
            grads_s = torch.autograd.grad(D_out4, leaves, g_logit, allow_unused=True)

        # ---- branch C (stream s2) — P4: discriminators on the target domain (label 0)
        with torch.cuda.stream(s2):
            with deferred_bn_updates(sink_p4), share_forward(memo):
                D_out4, leaves = self._d4_fresh(I["vertT"].transpose(2, 1))
            loss_D_diff_4, acc_t, g_logit = bce_with_logits_grad(D_out4, 0.0)
            grads_t = torch.autograd.grad(D_out4, leaves, g_logit, allow_unused=True)

        # ---- join: parameter gradients into the flat bucket, BatchNorm buffers in the reference's pass order
        if self.concurrent:
            for s in (s1, s2, s3):
                main.wait_stream(s)
            if not torch.cuda.is_current_stream_capturing():
                # eager mode: tensors born on a side stream are consumed on this one after the join
                crossing = list(grads_s) + list(grads_t) + [loss_D_same4, loss_D_diff_4, acc_s, acc_t]
                crossing += [loss_seg3, loss_vert_target, ent_mean]
                crossing += [t for sink in (sink_p3, sink_p4) for e in sink for t in (e[1], e[2])]
                for t in crossing:
                    if t is not None:
                        t.record_stream(main)
        # bucket = (grads_s + grads_t) / world: one launch; the division is the "mean" of the gradient all-reduce
        self._pending = None
        if any(g is None for g in grads_s) or any(g is None for g in grads_t):
            self.bucket.assign_sum(grads_s, grads_t)          # a parameter without gradient: zero-fill + accumulate
            if self._world > 1:
                self.bucket.flat.div_(self._world)
        elif self._world == 1:
            self._pending = (grads_s, grads_t)                # nothing to exchange: summed inside the SGD launch (_post)
        else:
            self.opt.pack(grads_s, grads_t, self.bucket.flat, scale=1.0 / self._world)
        apply_bn_updates([sink_p2, sink_p3, sink_p4])
        self._res = [loss_seg3.detach(), loss_vert_target.detach(), ent_mean.detach(), loss_adv_point.detach(),
                     loss_D_same4.detach(), loss_D_diff_4.detach(), acc_s, acc_t]

    def _exchange(self) -> None:
        """Sum of the (pre-divided) gradient bucket over the ranks; sets where SGD reads the result."""
        if self._world == 1:                 # single process, or exchange="local" inside a multi-rank job
            self._grad_final = self.bucket.flat
            return
        if self.exchange == "p2p":
            self._grad_final = self.comm.allreduce_p2p(self.bucket.numel)
        elif self.exchange == "nccl":
            self.comm.allreduce_(self.bucket.flat)
            self._grad_final = self.bucket.flat
        else:
            self.bucket.allreduce_sum()
            self._grad_final = self.bucket.flat

    def _post(self) -> None:
        # P5 (after the gradient all-reduce): SGD step of D4, scalars of the step into one small tensor
        main = torch.cuda.current_stream(self.device)
        side = self._side[2] if self.concurrent else main
        if self.concurrent:
            # the scalars of the step are gathered beside the optimiser launch, not after it
            side.wait_stream(main)
            if not torch.cuda.is_current_stream_capturing():
                for r in self._res:
                    r.record_stream(side)
        with torch.cuda.stream(side):
            self.result.copy_(torch.stack([r.reshape(()).float() for r in self._res]))
        if self._pending is not None:
            self.opt.sum_step(self._pending[0], self._pending[1], self.bucket.flat)
            self._pending = None
        else:
            self.opt.step(self._grad_final)
        if self.concurrent:
            main.wait_stream(side)

    def _sync_world(self) -> None:
        import torch.distributed as dist
        self._world = dist.get_world_size() if dist.is_available() and dist.is_initialized() and not self._force_local else 1

    def _body(self) -> None:
        self._sync_world()
        self._pre()
        self._exchange()                 # the bucket was filled pre-divided by the world size
        self._post()

    def run(self) -> torch.Tensor:
        """One step on the current stream (eager launches or graph replay). Returns the device-side
        result vector (RESULT_FIELDS)."""
        if self.graph is not None:
            self.graph.replay()
            if self.graph_post is not None:
                # exchange="torch": the one torch.distributed call of the step sits between the two captured halves
                self._exchange()
                self.graph_post.replay()
        else:
            self._body()
        return self.result

    def close(self) -> None:
        """Release the captured graphs, then the communicator (collective: every rank calls it).  The order matters:
        NCCL keeps a communicator alive while a CUDA graph that captured one of its collectives exists, so
        ncclCommDestroy waits for the graph's destruction."""
        import gc
        self.graph = self.graph_post = None
        gc.collect()
        torch.cuda.synchronize(self.device)
        if self.comm is not None:
            self.comm.check_status()
            self.comm.destroy()
            self.comm = None

    def _snapshot_state(self):
        """Everything a step mutates besides its outputs: D4's parameters and buffers (BatchNorm running statistics,
        num_batches_tracked), the optimiser's momentum and the device RNG (Dropout)."""
        return ([p.detach().clone() for p in self.bucket.params],
                [b.detach().clone() for b in self.d4.buffers()],
                self.opt.momentum_buffer.clone(), torch.cuda.get_rng_state(self.device))

    def _restore_state(self, snap) -> None:
        params, bufs, mom, rng = snap
        with torch.no_grad():
            for p, q in zip(self.bucket.params, params):
                p.copy_(q)
            for b, q in zip(self.d4.buffers(), bufs):
                b.copy_(q)
            self.opt.momentum_buffer.copy_(mom)
        torch.cuda.set_rng_state(rng, self.device)

    def capture(self, warmup: int = 3) -> None:
        """Capture the step into ONE CUDA graph (single GPU, or the gradient exchange done by libpcuda's own
        communicator: exchange = "p2p" / "nccl").  Only exchange="torch" is captured as two graphs around an
        ordinary torch.distributed call.

        The warm-up steps that precede the capture (allocator pools, lazy library state) run for real on whatever
        the input buffers hold; the training state they touch -- D4's parameters, BatchNorm buffers and
        num_batches_tracked, the SGD momentum, the Dropout RNG -- is saved before and restored after, so capturing
        a model mid-training or right after load_state_dict() does not move it off the reference trajectory."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and self.exchange == "torch"
        self._sync_world()
        torch.cuda.synchronize(self.device)
        snap = self._snapshot_state()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._restore_state(snap)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        if not multi:
            with torch.cuda.graph(g):
                self._body()
            self.graph, self.graph_post = g, None
            return
        pool = torch.cuda.graph_pool_handle()
        with torch.cuda.graph(g, pool=pool):
            self._pre()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=pool):
            self._post()
        self.graph, self.graph_post = g, g2

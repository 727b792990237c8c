"""Host-side logic of the batch-sharded step on CPU: two gloo ranks.

Only the plumbing is exercised here (sharding arithmetic, the flat gradient bucket and its
all-reduce); the kernels themselves need a GPU and are covered by `-m gpu` tests.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointcloududa_b200 import dist as pdist


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_the_batch():
    for batch in (8, 32, 33, 256, 5):
        for world in (1, 2, 4, 8):
            if batch < world:
                continue
            rs = [pdist.shard_range(batch, r, world) for r in range(world)]
            flat = [i for r in rs for i in r]
            assert flat == list(range(batch))
            assert max(len(r) for r in rs) - min(len(r) for r in rs) <= 1
    with pytest.raises(ValueError):
        pdist.shard_range(8, 2, 2)
    with pytest.raises(ValueError):
        pdist.check_per_rank_batch(8, 8)      # one cloud per rank: D4 cannot run (reference B == 1 branch)
    pdist.check_per_rank_batch(8, 4)


def test_grad_bucket_single_process():
    lin = torch.nn.Linear(3, 2)
    bucket = pdist.GradBucket(lin.parameters())
    assert bucket.numel == 8
    bucket.accumulate([torch.ones(2, 3), None])
    bucket.accumulate([torch.ones(2, 3), torch.full((2,), 3.0)])
    bucket.attach()
    assert torch.equal(lin.weight.grad, torch.full((2, 3), 2.0)) and torch.equal(lin.bias.grad, torch.full((2,), 3.0))
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    assert bucket.allreduce_mean() is None        # no process group: a no-op
    bucket.zero()
    assert float(lin.weight.grad.abs().sum()) == 0.0


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(world)
    os.environ["LOCAL_RANK"] = str(rank)
    try:
        r, _, w = pdist.init_from_env("gloo")
        assert (r, w) == (rank, world)
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 1))   # same init on both ranks
        bucket = pdist.GradBucket(net.parameters())
        # batch-sharded "step": each rank sees its shard of one global batch
        g = torch.Generator().manual_seed(5)
        xb = torch.randn(8, 3, generator=g)
        xs = pdist.shard_batch(xb, rank, world)
        loss = net(xs).pow(2).mean()
        bucket.zero()
        bucket.accumulate(torch.autograd.grad(loss, bucket.params))
        bucket.attach()
        bucket.allreduce_mean()
        # reference: the same loss on the whole batch in one process
        ref = torch.autograd.grad(net(xb).pow(2).mean(), list(net.parameters()))
        ref_flat = torch.cat([t.reshape(-1) for t in ref])
        err = float((bucket.flat - ref_flat).abs().max())
        # the step's variant: the bucket is filled pre-divided by the world size, the collective is a plain sum
        bucket.zero()
        bucket.accumulate(torch.autograd.grad(net(xs).pow(2).mean(), bucket.params))
        bucket.flat.div_(world)
        bucket.allreduce_sum()
        err = max(err, float((bucket.flat - ref_flat).abs().max()))
        sc = pdist.allreduce_scalars([torch.tensor(float(rank)), torch.tensor(2.0)])
        q.put((rank, err, sc.tolist()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e), None))


def test_two_rank_gloo_bucket_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, sc in got:
        assert isinstance(err, float), f"rank {rank} failed: {err}"
        assert err < 1e-6                       # mean of shard gradients == full-batch gradient
        assert sc == [0.5, 2.0]

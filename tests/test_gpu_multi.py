"""Two-rank tests of the multi-GPU path on real GPUs (skipped with fewer than 2 devices): libpcuda's communicator
(`pcuda_comm_*`: NCCL all-reduce and the NVLink peer-memory all-reduce kernel) and the batch-sharded
AdversarialStep with every exchange mode, eager and as ONE captured CUDA graph.

Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the log of the last such run is
committed under profiles/.  The CPU-side logic of the sharding is covered by tests/test_dist_gloo.py (gloo).
"""
import os
import socket
import sys
import traceback

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_main(fn_name, rank, world, port):
    """Entry of one rank (a fresh interpreter: `python tests/test_gpu_multi.py <fn> <rank> <world> <port>`)."""
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)        # a deadlocked rank prints where it is stuck and dies
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        globals()[fn_name](rank, world)
    except BaseException:
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)                                           # no collective on the way out: the peer must not wait for us
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: ok", flush=True)


def _run(fn_name, world=2, timeout=240):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import subprocess
    import tempfile
    import time
    port = _free_port()
    logs = [tempfile.NamedTemporaryFile("w+", suffix=f".rank{r}.log", delete=False) for r in range(world)]
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), fn_name, str(r), str(world), str(port)],
                              stdout=logs[r], stderr=subprocess.STDOUT, cwd=ROOT) for r in range(world)]
    t0 = time.time()
    failed = False
    while any(p.poll() is None for p in procs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() - t0 > timeout:
            failed = True
            time.sleep(2.0)
            break
        time.sleep(0.2)
    for p in procs:
        if p.poll() is None:
            p.kill()
    out = []
    for r, f in enumerate(logs):
        f.flush(); f.seek(0)
        out.append(f"---- rank {r} (exit {procs[r].poll()}) ----\n" + f.read()[-6000:])
        f.close(); os.unlink(f.name)
    assert not failed and all(p.returncode == 0 for p in procs), "\n".join(out)


# ---------------------------------------------------------------------------------------------------------
def _comm_body(rank, world):
    import torch.distributed as dist
    from pointcloududa_b200 import dist as pdist
    dev = torch.device("cuda", rank)
    n = 1_600_013                                   # odd on purpose: slices are padded to whole float4 per rank
    comm = pdist.PcudaComm(dev, p2p_floats=n)
    assert comm.world == world and comm.rank == rank and comm.nccl_version > 20000
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(n, generator=g).to(dev)
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    want = gathered[0].double()
    for t in gathered[1:]:
        want = want + t.double()
    # NCCL through libpcuda's own communicator
    y = x.clone()
    comm.allreduce_(y)
    torch.cuda.synchronize()
    assert (y.double() - want).abs().max().item() <= 1e-6 * want.abs().max().item()
    assert comm.p2p, "peer memory should be available between two GPUs of one NVSwitch node"
    # peer-memory kernel: repeated launches (epoch flags), shorter counts, bit-identical on all ranks,
    # fixed rank-order sum
    exact = gathered[0].clone()
    for t in gathered[1:]:
        exact = exact + t
    for it, cnt in enumerate([n, n, 1000, 4, n, 12345, n]):
        comm.buf_in[:cnt].copy_(x[:cnt] * (it + 1))
        out = comm.allreduce_p2p(cnt).clone()
        torch.cuda.synchronize()
        ref = gathered[0][:cnt] * (it + 1)
        for t in gathered[1:]:
            ref = ref + t[:cnt] * (it + 1)
        assert torch.equal(out, ref), (it, cnt, (out - ref).abs().max().item())
        other = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(other, out)
        assert all(torch.equal(o, out) for o in other)
    # the kernel inside a CUDA graph, replayed
    comm.buf_in[:n].copy_(x)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        comm.allreduce_p2p(n)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        out = comm.allreduce_p2p(n)
    for _ in range(5):
        gph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, exact)
    comm.check_status()
    dist.barrier()
    comm.destroy()


def test_comm_nccl_and_peer_memory_allreduce():
    _run("_comm_body")


# ---------------------------------------------------------------------------------------------------------
def _step_body(rank, world):
    import torch.distributed as dist
    from oracle import torch_step
    from pointcloududa_b200.step import AdversarialStep, StepConfig
    dev = torch.device("cuda", rank)
    w = dict(B=4, C=4, H=32, W=32, N=300, activation="sigmoid", normalize=False, return_prob=False)
    host = torch_step.conditioned_inputs(w, seed=500 + rank)          # each rank its own shard of the global batch
    ref_params = None
    for exchange in ("p2p", "nccl", "torch"):
        for graph in (False, True):
            cfg = StepConfig(B=w["B"], C=w["C"], H=w["H"], W=w["W"], N=w["N"], precision="fp32", lr_dis=2.5e-3)
            st = AdversarialStep(cfg, dev, seed=0, exchange=exchange)
            assert st.exchange == exchange and st._world == world
            for m in st.d4.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            st.load_inputs(host, non_blocking=False)
            if graph:
                st.capture(warmup=1)
                assert (st.graph_post is None) == (exchange != "torch")     # p2p / nccl: ONE graph holds the exchange
            for it in range(2):
                st.run()
                torch.cuda.synchronize()
                # the exchanged bucket == sum over ranks of the per-rank (pre-divided) gradients == mean gradient
                if exchange == "p2p":
                    local = st.bucket.flat.clone()                            # the peer-memory input buffer is left intact
                    parts = [torch.empty_like(local) for _ in range(world)]
                    dist.all_gather(parts, local)
                    want = parts[0].clone()
                    for t in parts[1:]:
                        want = want + t
                    assert torch.equal(st._grad_final, want), (exchange, graph, it)
            params = torch.cat([p.detach().reshape(-1) for p in st.d4.parameters()])
            # every rank holds the same parameters after the two steps ...
            allp = [torch.empty_like(params) for _ in range(world)]
            dist.all_gather(allp, params)
            for t in allp:
                assert torch.equal(t, params) if exchange == "p2p" else (t - params).abs().max().item() <= 1e-7, (exchange, graph)
            # ... and every exchange mode / eager vs graph replay lands on the same parameters
            if ref_params is None:
                ref_params = params.clone()
            else:
                d = (params - ref_params).abs().max().item()
                assert d <= 2e-6 * ref_params.abs().max().item(), (exchange, graph, d)
            dist.barrier()
            st.close()            # graphs first, then the communicator (collective)
            print(f"rank {rank}: exchange={exchange} graph={graph} ok", flush=True)


def test_adversarial_step_two_ranks_every_exchange_mode():
    _run("_step_body")


if __name__ == "__main__":
    _worker_main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))

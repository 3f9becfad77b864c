"""CPU, world_size 2, gloo: the data-parallel host logic (gradient SUM all-reduce in bucket slices and the loss
scaling that makes the sum reproduce the reference's global-batch gradient)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from disentangle_mlp_b200.trainer import GradReducer

    red = GradReducer()
    assert red.on and red.world == world
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    for lo, hi in red.buckets(flat.numel(), bucket_elems=300):
        red.allreduce_async(flat, lo, hi)
    red.wait()
    ok = torch.equal(flat, torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world)))
    # BCE is a mean over the GLOBAL batch: local mean-gradient scaled by 1/world, then SUM == global mean-gradient
    torch.manual_seed(0)
    p_all = torch.rand(8) * 0.8 + 0.1
    shard = p_all[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
    loss = torch.nn.functional.binary_cross_entropy(shard, torch.full((4,), 0.9)) * red.bce_scale()
    loss.backward()
    g = torch.zeros(8)
    g[rank * 4:(rank + 1) * 4] = shard.grad
    dist.all_reduce(g)
    ref = p_all.clone().requires_grad_(True)
    torch.nn.functional.binary_cross_entropy(ref, torch.full((8,), 0.9)).backward()
    ok = ok and torch.allclose(g, ref.grad, rtol=1e-6, atol=1e-8)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_scaling():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]

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


def _worker_schedule(rank, world, port, q):
    """The trainers' all-reduce schedule of one phase (early big buckets in bf16, early decoder tail, phase-end rest)
    driven through the real GradReducer over gloo on a FlatParams shell with CPU buffers."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace

    from disentangle_mlp_b200 import trainer as tr

    # a scaled-down VAE-like parameter list (names matter: EARLY_BUCKETS / the decoder tail are found by name)
    named = [("features.0.weight", (64, 3, 5, 5)), ("features.0.bias", (64,)), ("features.3.weight", (128, 64, 5, 5)),
             ("x_to_mu.0.weight", (256, 1024)), ("x_to_mu.0.bias", (256,)), ("x_to_mu.3.weight", (128, 256)),
             ("x_to_logvar.0.weight", (256, 1024)), ("x_to_logvar.3.weight", (128, 256)),
             ("preprocess.0.weight", (1024, 128)), ("deconv1.weight", (64, 64, 5, 5)), ("deconv4.bias", (3,))]
    plan = tr.plan_layout(named, True)
    fp = object.__new__(tr.FlatParams)
    fp.names, fp.offsets, fp.total, fp._small_end = plan["names"], plan["offsets"], plan["total"], plan["small_end"]
    fp.off16, fp._late_ranges, fp._tail_done = plan["off16"], plan["late_ranges"], None
    fp.shard = False
    fp.P = {n: SimpleNamespace(numel=lambda k=plan["numel"][n]: k) for n, _ in named}
    g = torch.Generator().manual_seed(100 + rank)
    fp.grad = torch.randn(plan["total"], generator=g)
    fp.grad16 = torch.randn(plan["total16"], generator=g).bfloat16()
    mine32, mine16 = fp.grad.clone(), fp.grad16.clone()
    red = tr.GradReducer()
    ok = red.on
    try:
        for n in ("x_to_mu.0.weight", "x_to_logvar.0.weight"):
            fp.reduce_early(red, n)
        fp.reduce_from(red, "preprocess.0.weight")
        fp.reduce_rest(red)
        red.wait()
        # expected: element-wise sums over the ranks (every element reduced exactly once)
        all32 = [torch.zeros_like(mine32) for _ in range(world)]
        all16 = [torch.zeros(mine16.numel()) for _ in range(world)]
        dist.all_gather(all32, mine32)
        dist.all_gather(all16, mine16.float())
        exp32 = sum(all32)
        for n, _ in named:
            o, k = plan["offsets"][n], plan["numel"][n]
            if n in plan["off16"]:
                o16 = plan["off16"][n]
                want = sum(a[o16:o16 + k] for a in all16)
                ok = ok and torch.allclose(fp.grad16[o16:o16 + k].float(), want, rtol=2e-2, atol=2e-2)
                ok = ok and torch.equal(fp.grad[o:o + k], mine32[o:o + k])  # its fp32 slot is not touched
            else:
                ok = ok and torch.allclose(fp.grad[o:o + k], exp32[o:o + k], rtol=1e-6, atol=1e-6)
    except Exception as e:  # noqa: BLE001
        print("schedule worker failed:", repr(e), flush=True)
        ok = False
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_sharded(rank, world, port, q):
    """Sharded optimizer step of a big tensor (ZeRO-1 on the 16384x2048 Linear weights): reduce-scatter of the bf16
    gradient, an Adam-like update of THIS rank's chunk only, all-gather of the bf16 shadow, and gather_masters() --
    through the real GradReducer / FlatParams methods over gloo.  Every rank must end up with the shadow (and, after
    gather_masters, the masters) that a replicated update on the summed gradient produces."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace

    from disentangle_mlp_b200 import trainer as tr

    ok = True
    try:
        n = 4096 * world
        red = tr.GradReducer()
        fp = object.__new__(tr.FlatParams)
        fp.reducer, fp.shard, fp._gather_pending = red, True, False
        fp.big16, fp.off16, fp.offsets = ["w"], {"w": 0}, {"w": 64}
        fp.P = {"w": SimpleNamespace(numel=lambda: n)}
        torch.manual_seed(7)
        init = torch.randn(64 + n)
        fp.flat, fp.m, fp.v = init.clone(), torch.zeros(64 + n), torch.zeros(64 + n)
        fp.shadow = init.bfloat16()
        g_all = [torch.randn(n, generator=torch.Generator().manual_seed(50 + r)).bfloat16() for r in range(world)]
        fp.grad16 = g_all[rank].clone()
        fp.reduce_early(red, "w")  # reduce-scatter (gloo: all-reduce)
        red.wait()
        a, b = red.chunk(n)
        gsum = sum(g.float() for g in g_all)
        ok = ok and torch.allclose(fp.grad16[a:b].float(), gsum[a:b], rtol=2e-2, atol=2e-2)
        # "Adam" on the own chunk: p -= 0.1 * g ; shadow = bf16(p)
        fp.flat[64 + a:64 + b] -= 0.1 * fp.grad16[a:b].float()
        fp.m[64 + a:64 + b] = fp.grad16[a:b].float()
        fp.shadow[64 + a:64 + b] = fp.flat[64 + a:64 + b].bfloat16()
        stale = fp.flat.clone()
        fp._gather_pending = True
        fp.gather_if_pending()
        fp.wait_gathered()
        want = init.clone()
        want[64:] -= 0.1 * gsum
        ok = ok and torch.allclose(fp.shadow[64:].float(), want[64:], rtol=2e-2, atol=2e-2)
        ok = ok and torch.equal(fp.shadow[:64], init[:64].bfloat16())  # neighbours untouched
        # masters: other ranks' chunks are stale until gather_masters()
        if world > 1:
            oa = 64 + ((rank + 1) % world) * (n // world)
            ok = ok and torch.equal(stale[oa:oa + 8], init[oa:oa + 8])
        fp.gather_masters()
        ok = ok and torch.allclose(fp.flat[64:], want[64:], rtol=2e-2, atol=2e-2)
        ok = ok and torch.allclose(fp.m[64:], gsum, rtol=2e-2, atol=2e-2)
    except Exception as e:  # noqa: BLE001
        import traceback

        traceback.print_exc()
        print("sharded worker failed:", repr(e), flush=True)
        ok = False
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_sharded_update_of_a_big_tensor():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_gloo_world2_reduce_schedule_on_flat_buffers():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_schedule, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


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

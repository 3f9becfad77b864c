"""Data-parallel consistency check (run under torchrun, >= 2 ranks): every rank trains on a DIFFERENT shard; because
every gradient element is all-reduced exactly once before each Adam step, the replicas must stay BITWISE identical.
A range that is never reduced makes the replicas drift apart; this is what the check catches.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_dp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    b = 16
    ok = True
    for workload in ("betavaegan", "vae", "gan"):
        for graph in (False, True):
            torch.manual_seed(999)
            np.random.seed(999)
            opt = dm.default_opt()
            if workload == "betavaegan":
                a, d = dm.VAE(opt), dm.Discriminator_celeba(opt)
                a.apply(dm.weights_init); d.apply(dm.weights_init)
                T = tr.BetaVAEGANTrainer(a.cuda(), d.cuda(), beta=25.0, lr=1e-3)
            elif workload == "gan":
                a, d = dm.Generator_celeba(opt), dm.Discriminator_celeba(opt)
                a.apply(dm.weights_init); d.apply(dm.weights_init)
                T = tr.GANTrainer(a.cuda(), d.cuda(), lr=3e-4)
            else:
                a = dm.VAE(opt); a.apply(dm.weights_init)
                T = tr.VAETrainer(a.cuda(), lr=3e-4)
            if graph:
                T.enable_graph(b)
            torch.manual_seed(1234 + rank)  # different data, noise and eps on every rank
            x = (torch.rand(b, 3, 64, 64) * 2 - 1).cuda()
            for _ in range(3):
                T.step(x)
            T.sync(masters=True)  # sharded Adam: gather the other ranks' chunks of the big tensors before comparing
            worst = 0.0
            for fp in T.flat_params():
                ref = fp.flat.clone()
                dist.broadcast(ref, 0)
                diff = (fp.flat - ref).abs().max()
                dist.all_reduce(diff, op=dist.ReduceOp.MAX)
                worst = max(worst, float(diff))
            if rank == 0:
                print(f"{workload:10s} graph={int(graph)}  world={world}  max |param(rank r) - param(rank 0)| after 3 steps = {worst:.3e}")
            ok = ok and worst == 0.0
            del T
    if rank == 0:
        print("DP CONSISTENT" if ok else "DP MISMATCH")
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""N-rank parity against the ORACLE: every rank trains on its shard of one global batch with
the fused trainer (gradients SUM-reduced over NCCL, BatchNorm per rank); rank 0 then runs the oracle's restated
reference step (oracle/steps.py) on the CONCATENATED batch with per-shard BatchNorm (steps.PerShard) and compares
 * the step's losses (each rank's local values summed over ranks = the reference's global-batch values),
 * every parameter after the step (all three Adam updates), and rank 0's BatchNorm running statistics.
Run under torchrun with >= 2 ranks (tests/test_ddp_gpu.py spawns it when the box has >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_dp_oracle.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr
from oracle import nets, steps  # test infrastructure: the checker, not the thing measured

SUM_KEYS = {"betavaegan": ["errD_real", "errD_fake", "errG_fake", "errG_recon", "sim", "recon_dec", "kld", "recon_enc"],
            "gan": ["errD", "errG"], "vae": ["loss"]}
# first-step tolerances (relative), as in tests/test_steps_gpu.py; parameters: relative L2 over all elements
TOL = {"errD_real": 5e-3, "errD_fake": 5e-3, "errG_fake": 2e-2, "errG_recon": 2e-2, "sim": 5e-2, "recon_dec": 2e-2,
       "kld": 0.15, "recon_enc": 5e-2, "errD": 5e-3, "errG": 3e-2, "loss": 1e-2}


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    b = int(os.environ.get("B", "16"))  # per rank
    nsteps = 2
    ok, report = True, {}
    for workload in ("betavaegan", "gan", "vae"):
        for graph in (False, True):
            torch.manual_seed(999)
            opt = steps.make_opt()
            if workload == "betavaegan":
                ra, rd = nets.VAE(opt), nets.Discriminator_celeba(opt)
            elif workload == "gan":
                ra, rd = nets.Generator_celeba(opt), nets.Discriminator_celeba(opt)
            else:
                ra, rd = nets.VAE(opt), None
            ra.apply(nets.weights_init)
            if rd is not None:
                rd.apply(nets.weights_init)
            ma = (dm.VAE(opt) if workload != "gan" else dm.Generator_celeba(opt)).cuda()
            ma.load_state_dict(ra.state_dict())
            md = None
            if rd is not None:
                md = dm.Discriminator_celeba(opt).cuda()
                md.load_state_dict(rd.state_dict())
            # (beta-VAE-GAN: 1e-4 instead of the reference's hard-coded 1e-3, at which KL after the first sign-like Adam
            # step is chaotic in the oracle itself -- see tests/test_bench_config_gpu.py; lr is a scalar kernel argument)
            lr = 1e-4 if workload == "betavaegan" else 3e-4
            if workload == "betavaegan":
                T = tr.BetaVAEGANTrainer(ma, md, beta=25.0, lr=lr)
            elif workload == "gan":
                T = tr.GANTrainer(ma, md, lr=lr)
            else:
                T = tr.VAETrainer(ma, lr=lr)
            if graph:
                T.enable_graph(b)
            x = steps.synthetic_batch(b * world, 4242)
            xs = x[rank * b:(rank + 1) * b].cuda()
            mine = []
            rands_all = []
            for s in range(nsteps):
                g = torch.Generator().manual_seed(300 + s)
                rands = [torch.randn(b * world, 128, generator=g) for _ in range(T.n_rands)]
                rands_all.append(rands)
                loc = [r[rank * b:(rank + 1) * b].cuda() for r in rands]
                if workload == "vae":
                    m = T.step(xs, loc[0])
                else:
                    m = T.step(xs, 0.9, 0.1, *loc)
                vals = torch.stack([m[k].float() for k in SUM_KEYS[workload]])
                dist.all_reduce(vals)  # local sums / (local mean / world) add up to the global-batch value
                mine.append({k: float(v) for k, v in zip(SUM_KEYS[workload], vals)})
            T.sync(masters=True)  # sharded Adam: fp32 masters of the big Linear weights complete on every rank
            torch.cuda.synchronize()
            if rank == 0:
                oa = torch.optim.Adam(ra.parameters(), lr=lr)
                od = torch.optim.Adam(rd.parameters(), lr=lr) if rd is not None else None
                pa = steps.PerShard(ra, world)
                pd = steps.PerShard(rd, world) if rd is not None else None
                tag = f"{workload}/graph={int(graph)}"
                rep = {}
                for s in range(nsteps):
                    if workload == "betavaegan":
                        ref = steps.betavaegan_step(pa, pd, oa, od, x, 25.0, 0.9, 0.1, *rands_all[s])
                    elif workload == "gan":
                        ref = steps.gan_step(pa, pd, oa, od, x, 0.9, 0.1, *rands_all[s])
                    else:
                        ref = steps.vae_step(pa, oa, x, *rands_all[s])
                    for k in SUM_KEYS[workload]:
                        r = abs(mine[s][k] - ref[k]) / (abs(ref[k]) + 1e-12)
                        rep[f"step{s}.{k}"] = r
                        # step 0 is the one-step statement; later steps inherit two bf16 Adam updates
                        if s == 0 and not r <= TOL[k]:
                            ok = False
                            print(f"MISMATCH {tag} step{s} {k}: cuda {mine[s][k]:.6g} oracle {ref[k]:.6g} rel {r:.3e}")
                pairs = [(ma, ra)] + ([(md, rd)] if rd is not None else [])
                for mm, rr in pairs:
                    a = torch.cat([p.detach().flatten().cpu() for p in mm.parameters()])
                    c = torch.cat([p.detach().flatten() for p in rr.parameters()])
                    name = type(rr).__name__
                    rep[f"params.{name}"] = rel(a, c)
                    if not rep[f"params.{name}"] < 0.1:
                        ok = False
                        print(f"MISMATCH {tag} params {name}: rel {rep['params.' + name]:.3e}")
                    sd_m, sd_r = mm.state_dict(), rr.state_dict()
                    worst = 0.0
                    for k, v in sd_r.items():
                        if "running" in k:
                            worst = max(worst, rel(sd_m[k].cpu(), v))
                        if "tracked" in k and int(sd_m[k]) != int(v):
                            ok = False
                            print(f"MISMATCH {tag} {k}: {int(sd_m[k])} vs {int(v)}")
                    rep[f"bn_running.{name}"] = worst
                    if not worst < 1e-1:
                        ok = False
                        print(f"MISMATCH {tag} BatchNorm running stats {name}: rel {worst:.3e}")
                report[tag] = rep
                print(tag, json.dumps({k: round(v, 5) for k, v in rep.items()}), flush=True)
            del T
            dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/dp_oracle_report.json", "w") as f:
            json.dump({"world": world, "per_rank_batch": b, "report": report, "ok": ok}, f, indent=1)
        print("DP ORACLE PARITY OK" if ok else "DP ORACLE PARITY FAILED", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()

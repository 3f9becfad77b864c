"""Per-launch timing of the GEMM-class kernel inside a real training step (warm caches, CUDA events on the
launching stream): python tools/gemm_breakdown.py [workload] [batch] -> gpurun_out/gemm_<workload>_b<batch>.csv"""
import csv
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    from disentangle_mlp_b200 import _lib, ops
    from disentangle_mlp_b200 import model as dm
    from disentangle_mlp_b200 import trainer as tr

    workload = sys.argv[1] if len(sys.argv) > 1 else "betavaegan"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.manual_seed(999)
    np.random.seed(999)
    opt = dm.default_opt()
    eg, d = dm.VAE(opt), dm.Discriminator_celeba(opt)
    eg.apply(dm.weights_init)
    d.apply(dm.weights_init)
    T = tr.BetaVAEGANTrainer(eg.cuda(), d.cuda(), beta=1.0, lr=1e-3)
    x = (torch.rand(b, 3, 64, 64) * 2 - 1).cuda()
    for _ in range(3):
        T.step(x)
    ops.profile_enable(True)
    ops.profile_read()
    nsteps = 5
    for _ in range(nsteps):
        T.step(x)
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/gemm_{workload}_b{b}.csv"
    _lib.check(_lib.load().dm_profile_dump(path.encode()), "dm_profile_dump")
    rows = list(csv.DictReader(open(path)))
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        k = (r.get("kind", "?") + r["mode"], r["a_mn"] + r["b_mn"], r["tiles_m"], r["tiles_n"], r["tiles_z"], r["bn"], r["kc"], r["stages"], r["ctas"])
        agg[k][0] += 1
        agg[k][1] += float(r["us"])
        agg[k][2] += float(r["gflop"])
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows) // nsteps} GEMM launches/step, {tot / nsteps / 1e3:.3f} ms/step")
    print("mode mn  tiles(m,n,z)      bn kc st ctas |  n/step   us/launch  ms/step   TFLOP/s")
    for k, (n, us, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[0]:>4} {k[1]:>2}  ({k[2]:>5},{k[3]:>4},{k[4]:>3}) {k[5]:>4} {k[6]:>2} {k[7]:>2} {k[8]:>4} | {n / nsteps:6.1f} {us / n:10.1f} {us / nsteps / 1e3:8.3f} {gf / us * 1e-3 * 1e6 / 1e3:9.1f}")


if __name__ == "__main__":
    main()

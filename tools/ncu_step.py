"""One profiled training step for ncu: warm up, cudaProfilerStart, ONE step, cudaProfilerStop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_step.py [betavaegan|gan|vae] [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "betavaegan"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.manual_seed(999)
    np.random.seed(999)
    opt = dm.default_opt()
    if workload == "betavaegan":
        eg, d = dm.VAE(opt), dm.Discriminator_celeba(opt)
        eg.apply(dm.weights_init); d.apply(dm.weights_init)
        T = tr.BetaVAEGANTrainer(eg.cuda(), d.cuda(), beta=1.0, lr=1e-3)
    elif workload == "gan":
        g, d = dm.Generator_celeba(opt), dm.Discriminator_celeba(opt)
        g.apply(dm.weights_init); d.apply(dm.weights_init)
        T = tr.GANTrainer(g.cuda(), d.cuda(), lr=3e-4)
    else:
        m = dm.VAE(opt); m.apply(dm.weights_init)
        T = tr.VAETrainer(m.cuda(), lr=3e-4)
    x = (torch.rand(b, 3, 64, 64) * 2 - 1).cuda()
    for _ in range(3):
        T.step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    T.step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one step", workload, b)


if __name__ == "__main__":
    main()

"""GPU probe: fused trainers vs the CPU oracle steps on identical weights / data / noise / labels."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr
from oracle import nets, steps


def rel(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def params_rel(mine, ref):
    a = torch.cat([p.detach().flatten().cpu() for p in mine.parameters()])
    b = torch.cat([p.detach().flatten() for p in ref.parameters()])
    return float((a - b).norm() / b.norm())


def main():
    b = int(os.environ.get("B", "16"))
    nsteps = int(os.environ.get("STEPS", "3"))
    which = os.environ.get("WHICH", "betavaegan,gan,vae").split(",")
    opt = steps.make_opt()
    x = steps.synthetic_batch(b, 1234)
    xc = x.cuda()
    if "betavaegan" in which:
        torch.manual_seed(999)
        rEG, rD = nets.VAE(opt), nets.Discriminator_celeba(opt)
        rEG.apply(nets.weights_init); rD.apply(nets.weights_init)
        mEG, mD = dm.VAE(opt).cuda(), dm.Discriminator_celeba(opt).cuda()
        mEG.load_state_dict(rEG.state_dict()); mD.load_state_dict(rD.state_dict())
        oEG = torch.optim.Adam(rEG.parameters(), lr=1e-3); oD = torch.optim.Adam(rD.parameters(), lr=1e-3)
        T = tr.BetaVAEGANTrainer(mEG, mD, beta=25.0, lr=1e-3)
        np.random.seed(999)
        for s in range(nsteps):
            real, fake = steps.draw_labels()
            g = torch.Generator().manual_seed(50 + s)
            noise, e1, e2 = (torch.randn(b, 128, generator=g) for _ in range(3))
            t0 = time.time()
            r = steps.betavaegan_step(rEG, rD, oEG, oD, x, 25.0, real, fake, noise, e1, e2)
            t1 = time.time()
            m = T.step(xc, real, fake, noise.cuda(), e1.cuda(), e2.cuda())
            m = {k: float(v) for k, v in m.items()}
            print(f"[bvg step {s}] cpu {t1 - t0:.2f}s labels {real},{fake}")
            for k in r:
                print(f"    {k:11s} ref {r[k]:14.6f} mine {m[k]:14.6f} rel {abs(m[k] - r[k]) / (abs(r[k]) + 1e-12):.3e}")
            print(f"    params rel: EG {params_rel(mEG, rEG):.3e} D {params_rel(mD, rD):.3e}")
    if "gan" in which:
        torch.manual_seed(999)
        rG, rD = nets.Generator_celeba(opt), nets.Discriminator_celeba(opt)
        rG.apply(nets.weights_init); rD.apply(nets.weights_init)
        mG, mD = dm.Generator_celeba(opt).cuda(), dm.Discriminator_celeba(opt).cuda()
        mG.load_state_dict(rG.state_dict()); mD.load_state_dict(rD.state_dict())
        oG = torch.optim.Adam(rG.parameters(), lr=3e-4); oD = torch.optim.Adam(rD.parameters(), lr=3e-4)
        T = tr.GANTrainer(mG, mD, lr=3e-4)
        np.random.seed(999)
        for s in range(nsteps):
            real, fake = steps.draw_labels()
            noise = torch.randn(b, 128, generator=torch.Generator().manual_seed(70 + s))
            r = steps.gan_step(rG, rD, oG, oD, x, real, fake, noise)
            m = {k: float(v) for k, v in T.step(xc, real, fake, noise.cuda()).items()}
            print(f"[gan step {s}]")
            for k in r:
                print(f"    {k:11s} ref {r[k]:14.6f} mine {m[k]:14.6f} rel {abs(m[k] - r[k]) / (abs(r[k]) + 1e-12):.3e}")
            print(f"    params rel: G {params_rel(mG, rG):.3e} D {params_rel(mD, rD):.3e}")
    if "vae" in which:
        torch.manual_seed(999)
        rV = nets.VAE(opt); rV.apply(nets.weights_init)
        mV = dm.VAE(opt).cuda(); mV.load_state_dict(rV.state_dict())
        oV = torch.optim.Adam(rV.parameters(), lr=3e-4)
        T = tr.VAETrainer(mV, lr=3e-4)
        for s in range(nsteps):
            eps = torch.randn(b, 128, generator=torch.Generator().manual_seed(90 + s))
            r = steps.vae_step(rV, oV, x, eps)
            m = {k: float(v) for k, v in T.step(xc, eps.cuda()).items()}
            print(f"[vae step {s}] loss ref {r['loss']:.4f} mine {m['loss']:.4f} rel {abs(m['loss'] - r['loss']) / r['loss']:.3e}"
                  f" params rel {params_rel(mV, rV):.3e}")
    from disentangle_mlp_b200 import _lib
    print("kernel launches:", _lib.launch_count())


if __name__ == "__main__":
    main()

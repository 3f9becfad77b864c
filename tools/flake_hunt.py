"""Run-to-run spread of the eager fused VAE step: many 5-step trials from the same initial state, relative deviation of
each step's loss from the median over trials.  Step 0 (initial parameters) agrees to ~1e-5; later steps differ by
1e-4 ... 1e-2 between IDENTICAL runs in every configuration (fp32 atomic order in split-K / weight-gradient / BatchNorm
reductions flips the sign of near-zero gradients under Adam's first, sign-like updates) -- the spread is the same with
PDL, side-stream weight gradients, fused heads or early Adam switched off, i.e. it is not a race.
    python tools/flake_hunt.py [trials]      env: the usual DM_* switches; HUNT_CFG=pdl0|wgs0|heads0 toggles in-process"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from disentangle_mlp_b200 import engine
from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr


def run(cfg, trials, b=16):
    os.environ["DM_PDL"] = "0" if "pdl0" in cfg else "1"
    engine.FUSE_HEADS = "heads0" not in cfg
    torch.manual_seed(999)
    opt = dm.default_opt()
    net = dm.VAE(opt)
    net.apply(dm.weights_init)
    net = net.cuda()
    T = tr.VAETrainer(net, lr=3e-4)
    if "wgs0" in cfg:
        engine.WgradSide.stream = None
    elif engine.WgradSide.stream is None:
        engine.WgradSide.stream = torch.cuda.Stream()
    if "early0" in cfg:
        T.BIG_ADAM = "now"
    snap = T.fp.snapshot()
    x = (torch.rand(b, 3, 64, 64, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    eps = [torch.randn(b, 128, generator=torch.Generator().manual_seed(90 + s)).cuda() for s in range(5)]
    out = []
    for t in range(trials):
        T.fp.restore(snap)
        out.append([float(T.step(x, eps[s])["loss"]) for s in range(5)])
    a = np.array(out)
    med = np.median(a, axis=0)
    dev = np.abs(a - med) / med
    print(f"cfg [{cfg or 'default'}] trials {trials}: per-step max rel dev {[f'{v:.1e}' for v in dev.max(axis=0)]}  "
          f"per-step median {[f'{v:.1e}' for v in np.median(dev, axis=0)]}", flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    for cfg in os.environ.get("HUNT_CFGS", "default,pdl0,wgs0,heads0,early0").split(","):
        run("" if cfg == "default" else cfg, n)

"""ORACLE — TEST INFRASTRUCTURE ONLY.

200-step loss curves of the oracle's three training steps (oracle/steps.py, which tests/test_oracle.py pins to the
real reference) at batch 16 with fully injected randomness, written to tests/golden/curves_*.json for the GPU
tracking test (tests/test_curves_gpu.py).  For each workload a second oracle run with the input images perturbed
by 1e-3 (relative) is recorded too: its deviation from the first run is the "chaos floor" -- how far two fp32
runs of the REFERENCE ITSELF drift apart under a perturbation of bf16-rounding size.

    python oracle/gen_curves.py [vae|gan|betavaegan ...]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nets, steps

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B, STEPS, SEED = 16, 200, 999


def rands(step, n):
    g = torch.Generator().manual_seed(10_000 + step)
    return [torch.randn(B, 128, generator=g) for _ in range(n)]


def labels():
    np.random.seed(SEED)
    return [steps.draw_labels() for _ in range(STEPS)]


def run(workload, x):
    torch.manual_seed(SEED)
    opt = steps.make_opt()
    lab = labels()
    out = []
    if workload == "vae":
        m = nets.VAE(opt)
        m.apply(nets.weights_init)
        o = torch.optim.Adam(m.parameters(), lr=3e-4)
        for s in range(STEPS):
            out.append(steps.vae_step(m, o, x, rands(s, 1)[0]))
    elif workload == "gan":
        g, d = nets.Generator_celeba(opt), nets.Discriminator_celeba(opt)
        g.apply(nets.weights_init)
        d.apply(nets.weights_init)
        og, od = torch.optim.Adam(g.parameters(), lr=3e-4), torch.optim.Adam(d.parameters(), lr=3e-4)
        for s in range(STEPS):
            out.append(steps.gan_step(g, d, og, od, x, lab[s][0], lab[s][1], rands(s, 1)[0]))
    else:
        eg, d = nets.VAE(opt), nets.Discriminator_celeba(opt)
        eg.apply(nets.weights_init)
        d.apply(nets.weights_init)
        oeg, od = torch.optim.Adam(eg.parameters(), lr=1e-3), torch.optim.Adam(d.parameters(), lr=1e-3)
        for s in range(STEPS):
            n, e1, e2 = rands(s, 3)
            out.append(steps.betavaegan_step(eg, d, oeg, od, x, 25.0, lab[s][0], lab[s][1], n, e1, e2))
    return out


def main():
    torch.set_num_threads(8)
    x = steps.synthetic_batch(B, 1234)
    gp = torch.Generator().manual_seed(4242)
    xp = x * (1 + 1e-3 * torch.randn(x.shape, generator=gp))
    for w in (sys.argv[1:] or ["vae", "gan", "betavaegan"]):
        t0 = time.time()
        base = run(w, x)
        pert = run(w, xp)
        keys = list(base[0].keys())
        doc = {"meta": {"batch": B, "steps": STEPS, "seed": SEED, "data_seed": 1234, "torch": torch.__version__,
                        "labels": labels(), "perturbation": "x * (1 + 1e-3 * N(0,1)), generator seed 4242"},
               "curves": {k: [r[k] for r in base] for k in keys},
               "perturbed": {k: [r[k] for r in pert] for k in keys}}
        with open(os.path.join(OUT, f"curves_{w}.json"), "w") as f:
            json.dump(doc, f)
        print(w, "done in", round(time.time() - t0), "s", flush=True)


if __name__ == "__main__":
    main()

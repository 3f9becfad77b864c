"""ORACLE — TEST INFRASTRUCTURE ONLY.

200-step loss curves of the oracle's three training steps (oracle/steps.py, which tests/test_oracle.py pins to the
real reference) at batch 16 with fully injected randomness, written to tests/golden/curves_*.json for the GPU
tracking test (tests/test_curves_gpu.py).  For each workload a second oracle run with the input images perturbed
by 1e-3 (relative) is recorded too: its deviation from the first run is the "chaos floor" -- how far two fp32
runs of the REFERENCE ITSELF drift apart under a perturbation of bf16-rounding size.

A third run, "bf16_emulated", is the same oracle with bf16 STORAGE emulated inside its fp32 arithmetic: the output
of every Conv / ConvTranspose / Linear / BatchNorm is rounded to bf16 (the cast's backward rounds the activation
gradient too) and conv / Linear weights are rounded to bf16 on use (fp32 masters keep receiving the Adam update).
Its deviation from the fp32 run is the "precision floor": what ANY bf16-storage implementation of the reference's
arithmetic costs on these trajectories.

    python oracle/gen_curves.py [vae|gan|betavaegan ...]             # all three runs
    python oracle/gen_curves.py --emulate-only [workloads ...]       # add / refresh only "bf16_emulated"
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


class _RoundBf16(torch.nn.Module):
    def forward(self, w):
        return w.bfloat16().float()


def _emulate_bf16_storage(net):
    import torch.nn.utils.parametrize as P
    from torch import nn

    for m in list(net.modules()):
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.BatchNorm2d, nn.BatchNorm1d)):
            m.register_forward_hook(lambda mod, inp, out: out.bfloat16().float())
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            P.register_parametrization(m, "weight", _RoundBf16(), unsafe=True)


def run(workload, x, emulate=False):
    torch.manual_seed(SEED)
    opt = steps.make_opt()
    lab = labels()
    out = []
    prep = _emulate_bf16_storage if emulate else (lambda net: None)
    if workload == "vae":
        m = nets.VAE(opt)
        m.apply(nets.weights_init)
        prep(m)
        o = torch.optim.Adam(m.parameters(), lr=3e-4)
        for s in range(STEPS):
            out.append(steps.vae_step(m, o, x, rands(s, 1)[0]))
    elif workload == "gan":
        g, d = nets.Generator_celeba(opt), nets.Discriminator_celeba(opt)
        g.apply(nets.weights_init)
        d.apply(nets.weights_init)
        prep(g), prep(d)
        og, od = torch.optim.Adam(g.parameters(), lr=3e-4), torch.optim.Adam(d.parameters(), lr=3e-4)
        for s in range(STEPS):
            out.append(steps.gan_step(g, d, og, od, x, lab[s][0], lab[s][1], rands(s, 1)[0]))
    else:
        eg, d = nets.VAE(opt), nets.Discriminator_celeba(opt)
        eg.apply(nets.weights_init)
        d.apply(nets.weights_init)
        prep(eg), prep(d)
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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    for w in (args or ["vae", "gan", "betavaegan"]):
        t0 = time.time()
        if "--emulate-only" in sys.argv:
            path = os.path.join(OUT, f"curves_{w}.json")
            doc = json.load(open(path))
            emu = run(w, x, emulate=True)
            doc["bf16_emulated"] = {k: [r[k] for r in emu] for k in emu[0]}
            doc["meta"]["bf16_emulated"] = "oracle with bf16 storage of layer outputs, their gradients and conv/Linear weights"
            with open(path, "w") as f:
                json.dump(doc, f)
            print(w, "bf16_emulated done in", round(time.time() - t0), "s", flush=True)
            continue
        base = run(w, x)
        pert = run(w, xp)
        emu = run(w, x, emulate=True)
        keys = list(base[0].keys())
        doc = {"meta": {"batch": B, "steps": STEPS, "seed": SEED, "data_seed": 1234, "torch": torch.__version__,
                        "labels": labels(), "perturbation": "x * (1 + 1e-3 * N(0,1)), generator seed 4242"},
               "curves": {k: [r[k] for r in base] for k in keys},
               "perturbed": {k: [r[k] for r in pert] for k in keys},
               "bf16_emulated": {k: [r[k] for r in emu] for k in keys}}
        with open(os.path.join(OUT, f"curves_{w}.json"), "w") as f:
            json.dump(doc, f)
        print(w, "done in", round(time.time() - t0), "s", flush=True)


if __name__ == "__main__":
    main()

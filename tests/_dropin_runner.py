"""Helper process of tests/test_dropin_scripts.py (not a test module).

Executes an UNMODIFIED reference training script (/root/reference/experiments/<script>) with `dropin/` on sys.path in
place of the reference's `models/`: `from model import *` and `from helper_functions import *` resolve to this repo's
kernel-backed classes / sampling helpers.  Stand-ins exist only for what cannot exist offline (`dataset`, `fid`,
matplotlib / IPython stubs), exactly as in oracle/gen_golden.py.  The build container has no GPU and the product has
no CPU path, so the ONE thing mocked is the kernel layer: `_KernelBacked._run` (the single entry every network
forward goes through; plus the reparameterisation op) delegates to the oracle's torch.nn restatement sharing the SAME Parameter / buffer objects.
Everything above it -- construction from `opt`, weights_init via .apply, DataParallel(...).module, the attribute
assignments at new_betavaegan.py:132-180, zero_grad, the six .backward() calls, torch.optim.Adam over our
parameters, state_dict -- is the real drop-in code driven by the real script.  Prints one JSON line."""
import importlib.util
import json
import os
import sys
import tempfile
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from oracle import gen_golden as G  # stand-ins, seeds and digests shared with the golden generator
from oracle import nets

REF = G.REF


def share(ours, theirs):
    """make the oracle module compute on OUR Parameter / buffer objects"""
    mods = dict(theirs.named_modules())
    for name, m in ours.named_modules():
        o = mods[name]
        for k, p in m._parameters.items():
            if p is not None:
                o._parameters[k] = p
        for k, b in m._buffers.items():
            if b is not None:
                o._buffers[k] = b


def install_mock():
    from disentangle_mlp_b200 import model as dm

    twins = {}

    def twin(owner):
        if id(owner) not in twins:
            cls = getattr(nets, type(owner).__name__)
            rng = torch.get_rng_state()  # constructing the twin draws default inits: keep the script's RNG stream
            t = cls(G.SimpleNamespace(input_channels=3, n_hidden=128, n_z=[256, 8, 8]))
            torch.set_rng_state(rng)
            twins[id(owner)] = t
        share(owner, twins[id(owner)])  # (re-share every call: .to() / load_state_dict may have replaced objects)
        return twins[id(owner)]

    def _run(self, kind, x):
        t = twin(self)
        t.train(self.training)
        if kind == "disc":
            prob, feat = t(x)  # nets.Discriminator_celeba.forward squeezes; ours squeezes again (no-op)
            return prob, feat
        if kind == "enc":
            return t.encode(x)
        return t.decode(x) if hasattr(t, "decode") else t(x)

    class _Reparam:  # the second kernel-level entry: z = mu + eps * exp(0.5 logvar) (models/model.py:532-535)
        @staticmethod
        def apply(mu, logvar, eps):
            return mu + eps * torch.exp(0.5 * logvar)

    dm._KernelBacked._run = _run
    dm._ReparamFn = _Reparam
    return dm


def main():
    script = sys.argv[1]
    scratch = Path(tempfile.mkdtemp(prefix="dm_dropin_"))
    standins = {k: v for k, v in G.STANDINS.items() if k != "helper_functions.py"}  # ours comes from dropin/
    for rel, src in standins.items():
        p = scratch / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(textwrap.dedent(src))
    os.chdir(scratch)
    sys.path[:0] = [str(scratch), str(ROOT / "dropin"), str(REF / "utils")]
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    os.environ["DM_STEPS"] = "1"
    os.environ["DM_DATA_SEED"] = str(G.DATA_SEED)
    dm = install_mock()
    torch.set_num_threads(8)
    out = {"script": script}
    if script == "new_vae.py":
        m = G.load_script(script, scratch, ["--lr", "3e-4"])
        model = m.model.module
        assert isinstance(model, dm.VAE), type(model)
        torch.manual_seed(100)
        out["avg_loss"] = float(m.train(0))
        out["params"] = G.tensor_digest(torch.cat([p.flatten() for p in m.model.parameters()]))
    elif script == "new_gan.py":
        m = G.load_script(script, scratch, ["--lr", "3e-4"])
        unwrap = lambda n: getattr(n, "module", n)  # noqa: E731 - new_gan.py wraps in DataParallel only with GPUs (:52-57)
        assert isinstance(unwrap(m.netG), dm.Generator_celeba) and isinstance(unwrap(m.netD), dm.Discriminator_celeba)
        m.epoch = 0
        np.random.seed(G.SEED)
        torch.manual_seed(200)
        g, _d = m.train()
        out["avg_loss_G"] = float(g)
        out["paramsG"] = G.tensor_digest(torch.cat([p.flatten() for p in m.netG.parameters()]))
        out["paramsD"] = G.tensor_digest(torch.cat([p.flatten() for p in m.netD.parameters()]))
    else:
        m = G.load_script(script, scratch, ["--beta", "25"])
        assert isinstance(m.netEG.module, dm.VAE) and isinstance(m.netD.module, dm.Discriminator_celeba)
        np.random.seed(G.SEED)
        torch.manual_seed(300)
        enc, dec, dis, dx = m.train(0)
        out.update(enc=float(enc), dec=float(dec), dis=float(dis), Dx=float(dx))
        out["paramsEG"] = G.tensor_digest(torch.cat([p.flatten() for p in m.netEG.parameters()]))
        out["paramsD"] = G.tensor_digest(torch.cat([p.flatten() for p in m.netD.parameters()]))
        out["bn_tracked"] = {"D": int(m.netD.module.convs[1].num_batches_tracked),
                             "Enc": int(m.netEG.module.features[1].num_batches_tracked),
                             "Dec": int(m.netEG.module.act1[0].num_batches_tracked)}
        sd = m.netEG.module.state_dict()
        out["n_state_keys"] = len(sd)
    print("DROPIN_RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

"""ORACLE — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.json by running the REAL reference (read-only checkout at /root/reference):
  * models/model.py classes, imported as `model`, for module-level vectors, and
  * experiments/new_vae.py, new_gan.py, new_betavaegan.py, executed UNMODIFIED via importlib with stand-ins
    only for the three modules that cannot exist offline (`dataset`, `helper_functions`, `fid`) plus empty
    `matplotlib` / `IPython` stubs that new_gan.py imports.
The reference cannot travel to the GPU box, so the vectors are committed; tests/test_oracle.py replays the
same seeds through oracle/nets.py + oracle/steps.py and must reproduce them.

    python oracle/gen_golden.py            (run in the build container; needs /root/reference)
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import textwrap
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REF = Path(os.environ.get("DM_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
SEED, BATCH, STEPS, DATA_SEED = 999, 4, 3, 1234


def tensor_digest(t: torch.Tensor):
    t = t.detach().double().flatten()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def params_digest(module):
    return {k: tensor_digest(v) for k, v in module.state_dict().items()}


def grads_digest(module):
    return {k: tensor_digest(p.grad) for k, p in module.named_parameters()}


def sample(t: torch.Tensor, n=64):
    """A small deterministic subsample of a tensor (every k-th element)."""
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return [float(v) for v in f[::step][:n]]


def batch(seed=DATA_SEED, b=BATCH):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, 64, 64, generator=g) * 2 - 1


def module_vectors():
    sys.path.insert(0, str(REF / "models"))
    import model as ref  # the reference's models/model.py

    opt = SimpleNamespace(input_channels=3, n_hidden=128, n_z=[256, 8, 8])
    out = {}
    torch.manual_seed(SEED)
    vae = ref.VAE(opt)
    dis = ref.Discriminator_celeba(opt)
    gen = ref.Generator_celeba(opt)
    enc = ref.Encoder_celeba(opt)
    for m in (vae, dis, gen, enc):
        m.apply(ref.weights_init)
    out["init"] = {"VAE": params_digest(vae), "D": params_digest(dis), "G": params_digest(gen), "E": params_digest(enc)}
    x = batch()
    # VAE forward + backward of (sum recon + sum mu - sum logvar)
    torch.manual_seed(7)
    recon, mu, logvar = vae(x)
    (recon.sum() + mu.sum() - logvar.sum()).backward()
    out["VAE"] = {"recon": sample(recon), "mu": sample(mu), "logvar": sample(logvar), "grads": grads_digest(vae),
                  "state_after": params_digest(vae)}
    # Discriminator
    prob, feat = dis(x)
    (prob.sum() + 0.01 * feat.pow(2).sum()).backward()
    out["D"] = {"prob": sample(prob), "feat": sample(feat), "grads": grads_digest(dis), "state_after": params_digest(dis)}
    # Generator
    torch.manual_seed(8)
    code = torch.randn(BATCH, 128)
    img = gen(code)
    img.pow(2).sum().backward()
    out["G"] = {"img": sample(img), "grads": grads_digest(gen)}
    # Encoder_celeba (z, kld)
    torch.manual_seed(9)
    z, kld = enc(x)
    (z.sum() + kld.sum()).backward()
    out["E"] = {"z": sample(z), "kld": sample(kld), "grads": grads_digest(enc)}
    sys.path.pop(0)
    return out


STANDINS = {
    "dataset.py": """
        import torch
        class _DS:
            def __init__(self, n): self.n = n
            def __len__(self): return self.n
        class _Loader:
            def __init__(self, b, steps, seed):
                self.b, self.steps, self.seed = b, steps, seed
                self.dataset = _DS(b * steps)
            def __len__(self): return self.steps
            def __iter__(self):
                g = torch.Generator().manual_seed(self.seed)
                for _ in range(self.steps):
                    yield torch.rand(self.b, 3, 64, 64, generator=g) * 2 - 1, torch.zeros(self.b)
        def get_data_loader(opt):
            import os
            l = _Loader(opt.batch_size_train, int(os.environ.get("DM_STEPS", "1")), int(os.environ.get("DM_DATA_SEED", "1234")))
            return l, l, l
    """,
    "helper_functions.py": """
        def generate_fid_samples(*a, **k): pass
        def gen_reconstructions(*a, **k): pass
        def generate_samples(*a, **k): pass
    """,
    "fid.py": "def get_fid(*a, **k): return float('nan')\n",
    "matplotlib/__init__.py": "", "matplotlib/pyplot.py": "", "matplotlib/animation.py": "",
    "IPython/__init__.py": "", "IPython/display.py": "HTML = None\n",
}


def load_script(name, scratch, extra_args=()):
    """exec the reference training script's module-level code (not its __main__ block)."""
    sys.argv = [name, "--name", "golden", "--batch_size_train", str(BATCH), "--calc_fid", "false", "--num_workers", "0",
                *extra_args]
    spec = importlib.util.spec_from_file_location("ref_" + name.replace(".py", ""), REF / "experiments" / name)
    mod = importlib.util.module_from_spec(spec)
    torch.manual_seed(SEED)  # new_gan.py builds its nets before seeding (new_gan.py:47-57 vs :155-156)
    spec.loader.exec_module(mod)
    return mod


def loop_vectors():
    out = {}
    scratch = Path(tempfile.mkdtemp(prefix="dm_golden_"))
    for rel, src in STANDINS.items():
        p = scratch / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(textwrap.dedent(src))
    cwd = os.getcwd()
    os.chdir(scratch)  # EnvSetter creates ./data/<name>/... in the cwd (utils/envsetter.py:23-24,68-87)
    sys.path[:0] = [str(scratch), str(REF / "models"), str(REF / "utils")]
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    os.environ["DM_STEPS"] = "1"
    try:
        # ---- VAE (new_vae.py): lr from --lr
        os.environ["DM_DATA_SEED"] = str(DATA_SEED)
        m = load_script("new_vae.py", scratch, ["--lr", "3e-4"])
        steps = []
        for s in range(STEPS):
            torch.manual_seed(100 + s)
            avg = m.train(0)
            steps.append({"avg_loss": float(avg), "params": tensor_digest(torch.cat([p.flatten() for p in m.model.parameters()]))})
        out["vae"] = steps
        # ---- GAN (new_gan.py)
        m = load_script("new_gan.py", scratch, ["--lr", "3e-4"])
        m.epoch = 0
        np.random.seed(SEED)
        steps = []
        for s in range(STEPS):
            torch.manual_seed(200 + s)
            g, d = m.train()
            steps.append({"avg_loss_G": float(g),
                          "paramsG": tensor_digest(torch.cat([p.flatten() for p in m.netG.parameters()])),
                          "paramsD": tensor_digest(torch.cat([p.flatten() for p in m.netD.parameters()]))})
        out["gan"] = steps
        # ---- beta-VAE-GAN (new_betavaegan.py): lr hard-coded 1e-3 (:49-50)
        m = load_script("new_betavaegan.py", scratch, ["--beta", "25"])
        np.random.seed(SEED)
        steps = []
        for s in range(STEPS):
            torch.manual_seed(300 + s)
            enc, dec, dis, dx = m.train(0)
            steps.append({"enc": float(enc), "dec": float(dec), "dis": float(dis), "Dx": float(dx),
                          "paramsEG": tensor_digest(torch.cat([p.flatten() for p in m.netEG.parameters()])),
                          "paramsD": tensor_digest(torch.cat([p.flatten() for p in m.netD.parameters()])),
                          "bn_tracked": {"D": int(m.netD.module.convs[1].num_batches_tracked),
                                         "Enc": int(m.netEG.module.features[1].num_batches_tracked),
                                         "Dec": int(m.netEG.module.act1[0].num_batches_tracked)}})
        out["betavaegan"] = steps
    finally:
        os.chdir(cwd)
        del sys.path[:3]
    return out


def main():
    torch.set_num_threads(int(os.environ.get("DM_GOLDEN_THREADS", "8")))
    OUT.mkdir(parents=True, exist_ok=True)
    meta = {"torch": torch.__version__, "seed": SEED, "batch": BATCH, "steps": STEPS, "data_seed": DATA_SEED,
            "reference": str(REF)}
    mv = module_vectors()
    (OUT / "modules.json").write_text(json.dumps({"meta": meta, **mv}))
    lv = loop_vectors()
    (OUT / "loops.json").write_text(json.dumps({"meta": meta, **lv}, indent=1))
    print("wrote", OUT / "modules.json", (OUT / "modules.json").stat().st_size, "bytes;",
          OUT / "loops.json", (OUT / "loops.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()

"""The oracle (oracle/nets.py, oracle/steps.py) against the golden vectors recorded from the REAL reference
(tests/golden/*.json, written by oracle/gen_golden.py).  CPU only.

Tolerance: the vectors were produced with 8 oneDNN threads on the build container; another host may pick
different CPU kernels / reduction orders, so comparisons use rtol 2e-4 on digests (they are bit-exact here).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nets, steps

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 2e-4


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def digest(t):
    t = t.detach().double().flatten()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def sample(t, n=64):
    f = t.detach().flatten()
    step = max(1, f.numel() // n)
    return [float(v) for v in f[::step][:n]]


def close(a, b, rtol=RTOL, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(1.0, float(np.abs(b).max()))
    assert a.shape == b.shape
    assert np.all(np.abs(a - b) <= atol * scale + rtol * np.maximum(np.abs(b), scale * 1e-3)), (a, b)


def digests_close(got: dict, want: dict, rtol=RTOL):
    assert list(got.keys()) == list(want.keys())
    for k in want:
        g, w = got[k], want[k]
        # the L1 and L2 digests are well conditioned; the plain sum can cancel, so scale by the L1 digest
        assert abs(g[1] - w[1]) <= rtol * max(w[1], 1e-12) + 1e-9, (k, g, w)
        assert abs(g[2] - w[2]) <= 2 * rtol * max(w[2], 1e-12) + 1e-12, (k, g, w)
        assert abs(g[0] - w[0]) <= rtol * max(w[1], 1e-12) + 1e-9, (k, g, w)


@pytest.fixture(scope="module")
def mods():
    g = _load("modules.json")
    torch.manual_seed(g["meta"]["seed"])
    opt = steps.make_opt()
    vae, dis, gen, enc = nets.VAE(opt), nets.Discriminator_celeba(opt), nets.Generator_celeba(opt), nets.Encoder_celeba(opt)
    for m in (vae, dis, gen, enc):
        m.apply(nets.weights_init)
    return g, vae, dis, gen, enc


def test_state_dict_keys_and_init(mods):
    g, vae, dis, gen, enc = mods
    digests_close({k: digest(v) for k, v in vae.state_dict().items()}, g["init"]["VAE"])
    digests_close({k: digest(v) for k, v in dis.state_dict().items()}, g["init"]["D"])
    digests_close({k: digest(v) for k, v in gen.state_dict().items()}, g["init"]["G"])
    digests_close({k: digest(v) for k, v in enc.state_dict().items()}, g["init"]["E"])
    assert len(vae.state_dict()) == 69 and len(dis.state_dict()) == 32 and len(gen.state_dict()) == 30


def test_module_forward_backward(mods):
    g, vae, dis, gen, enc = mods
    b = g["meta"]["batch"]
    x = steps.synthetic_batch(b, g["meta"]["data_seed"])
    torch.manual_seed(7)
    recon, mu, logvar = vae(x)
    (recon.sum() + mu.sum() - logvar.sum()).backward()
    close(sample(recon), g["VAE"]["recon"])
    close(sample(mu), g["VAE"]["mu"])
    close(sample(logvar), g["VAE"]["logvar"])
    digests_close({k: digest(p.grad) for k, p in vae.named_parameters()}, g["VAE"]["grads"], rtol=1e-3)
    digests_close({k: digest(v) for k, v in vae.state_dict().items()}, g["VAE"]["state_after"])
    prob, feat = dis(x)
    (prob.sum() + 0.01 * feat.pow(2).sum()).backward()
    close(sample(prob), g["D"]["prob"])
    close(sample(feat), g["D"]["feat"])
    digests_close({k: digest(p.grad) for k, p in dis.named_parameters()}, g["D"]["grads"], rtol=1e-3)
    torch.manual_seed(8)
    img = gen(torch.randn(b, 128))
    img.pow(2).sum().backward()
    close(sample(img), g["G"]["img"])
    digests_close({k: digest(p.grad) for k, p in gen.named_parameters()}, g["G"]["grads"], rtol=1e-3)
    torch.manual_seed(9)
    z, kld = enc(x)
    (z.sum() + kld.sum()).backward()
    close(sample(z), g["E"]["z"])
    close(sample(kld), g["E"]["kld"])


def _cat(m):
    return digest(torch.cat([p.flatten() for p in m.parameters()]))


def test_vae_loop():
    g = _load("loops.json")
    b, seed = g["meta"]["batch"], g["meta"]["seed"]
    torch.manual_seed(seed)
    model = nets.VAE(steps.make_opt())
    model.apply(nets.weights_init)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    x = steps.synthetic_batch(b, g["meta"]["data_seed"])
    for s, want in enumerate(g["vae"]):
        torch.manual_seed(100 + s)
        r = steps.vae_step(model, opt, x)
        close([r["loss"] / b], [want["avg_loss"]])
        close(_cat(model), want["params"], rtol=1e-4)


def test_gan_loop():
    g = _load("loops.json")
    b, seed = g["meta"]["batch"], g["meta"]["seed"]
    torch.manual_seed(seed)
    o = steps.make_opt()
    netG, netD = nets.Generator_celeba(o), nets.Discriminator_celeba(o)
    netG.apply(nets.weights_init)
    netD.apply(nets.weights_init)
    optG = torch.optim.Adam(netG.parameters(), lr=3e-4)
    optD = torch.optim.Adam(netD.parameters(), lr=3e-4)
    x = steps.synthetic_batch(b, g["meta"]["data_seed"])
    np.random.seed(seed)
    for s, want in enumerate(g["gan"]):
        torch.manual_seed(200 + s)
        torch.randn(b, 128)  # `fixed_noise`, drawn once per train() call and never used (new_gan.py:69)
        real, fake = steps.draw_labels()
        r = steps.gan_step(netG, netD, optG, optD, x, real, fake)
        close([r["errG"] / b], [want["avg_loss_G"]], rtol=2e-3)
        close(_cat(netG), want["paramsG"], rtol=1e-4)
        close(_cat(netD), want["paramsD"], rtol=1e-4)


def test_betavaegan_loop():
    g = _load("loops.json")
    b, seed = g["meta"]["batch"], g["meta"]["seed"]
    torch.manual_seed(seed)
    o = steps.make_opt()
    netEG, netD = nets.VAE(o), nets.Discriminator_celeba(o)
    netEG.apply(nets.weights_init)
    netD.apply(nets.weights_init)
    optEG = torch.optim.Adam(netEG.parameters(), lr=1e-3)
    optD = torch.optim.Adam(netD.parameters(), lr=1e-3)
    x = steps.synthetic_batch(b, g["meta"]["data_seed"])
    np.random.seed(seed)
    for s, want in enumerate(g["betavaegan"]):
        torch.manual_seed(300 + s)
        real, fake = steps.draw_labels()
        r = steps.betavaegan_step(netEG, netD, optEG, optD, x, 25.0, real, fake)
        close([r["recon_enc"] / b], [want["enc"]], rtol=1e-3)
        close([r["D_x"] / b], [want["Dx"]], rtol=1e-3)
        close(_cat(netEG), want["paramsEG"], rtol=1e-4)
        close(_cat(netD), want["paramsD"], rtol=1e-4)
        # BatchNorm running stats are updated on every forward: D 5x, encoder 2x, decoder 3x per step
        assert int(netD.convs[1].num_batches_tracked) == want["bn_tracked"]["D"] == 5 * (s + 1)
        assert int(netEG.features[1].num_batches_tracked) == want["bn_tracked"]["Enc"] == 2 * (s + 1)
        assert int(netEG.act1[0].num_batches_tracked) == want["bn_tracked"]["Dec"] == 3 * (s + 1)

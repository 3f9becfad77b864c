"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU fp32 restatement of ONE iteration of the three reference training loops, on stock torch autograd and
torch.optim.Adam.  Each function follows its loop body statement by statement (same number and order of
forward passes, `.backward()` calls and optimizer steps), so with the same seeds it reproduces the
reference bit for bit (tests/test_oracle.py checks this against tests/golden/).

  vae_step          experiments/new_vae.py:39-60
  gan_step          experiments/new_gan.py:74-128
  betavaegan_step   experiments/new_betavaegan.py:64-75 (losses), 87-193 (loop body)

Random draws: when `noise` / `eps*` are None they are drawn from torch's global generator in the
reference's order (new_betavaegan.py:111 randn(B,128); then one randn_like per VAE forward, model.py:534);
tests inject them instead so the CUDA path can be fed the same numbers.
Labels (`real_label`, `fake_label`) are per-step scalars drawn by the caller, as the reference draws them
from numpy's global RNG (new_betavaegan.py:89-90, new_gan.py:77-78): see `draw_labels`.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F


def draw_labels():
    """new_betavaegan.py:89-90 / new_gan.py:77-78 — fake first, then real, from numpy's global RNG."""
    fake = np.random.choice(a=[0.1, 0.9], p=[0.95, 0.05])
    real = np.random.choice(a=[0.1, 0.9], p=[0.05, 0.95])
    return float(real), float(fake)


def _vae_forward(model, x, eps):
    if eps is None:
        return model(x)
    mu, logvar = model.encode(x)
    return model.decode(mu + eps * torch.exp(0.5 * logvar)), mu, logvar


def kld_sum(mu, logvar):
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())


def vae_step(model, optimizer, data, eps=None):
    """new_vae.py:53-60: loss = MSE(sum) + KL(sum), beta = 1 (loss_function, :39-48)."""
    optimizer.zero_grad()
    recon, mu, logvar = _vae_forward(model, data, eps)
    loss = F.mse_loss(recon, data, reduction="sum") + kld_sum(mu, logvar)
    loss.backward()
    optimizer.step()
    return {"loss": loss.item()}


def gan_step(netG, netD, optimizerG, optimizerD, data, real_label, fake_label, noise=None):
    """new_gan.py:84-128."""
    criterion = nn.BCELoss()
    b = data.size(0)
    netD.zero_grad()
    label = torch.full((b,), real_label, device=data.device)
    output, _ = netD(data)
    errD_real = criterion(output, label)
    errD_real.backward()
    D_x = output.mean().item()
    if noise is None:
        noise = torch.randn(b, 128, device=data.device)
    fake = netG(noise)
    label.fill_(fake_label)
    output, _ = netD(fake.detach())
    errD_fake = criterion(output, label)
    errD_fake.backward()
    D_G_z1 = output.mean().item()
    optimizerD.step()

    netG.zero_grad()
    label.fill_(real_label)
    output, _ = netD(fake)
    errG = criterion(output, label)
    errG.backward()
    D_G_z2 = output.mean().item()
    optimizerG.step()
    return {"errD": (errD_real + errD_fake).item(), "errG": errG.item(), "D_x": D_x, "D_G_z1": D_G_z1,
            "D_G_z2": D_G_z2}


def betavaegan_step(netEG, netD, optimizerEG, optimizerD, data, beta, real_label, fake_label, noise=None,
                    eps_dec=None, eps_enc=None):
    """new_betavaegan.py:93-193.  The `module.requires_grad = ...` assignments (:132-143, :169-180) set a plain
    attribute on nn.Module objects and have no effect, so both EG updates touch every encoder and decoder
    parameter; `sim_real` is not detached (:129, :160)."""
    criterion = nn.BCELoss()
    b = data.size(0)
    # ---- discriminator (:95-123)
    netD.zero_grad()
    label = torch.full((b,), real_label, device=data.device)
    output, sim_real = netD(data)
    errD_real = criterion(output, label)
    errD_real.backward()
    D_x = output.mean().item()
    if noise is None:
        noise = torch.randn(b, 128, device=data.device)
    fake = netEG.decode(noise)
    label.fill_(fake_label)
    output, _ = netD(fake.detach())
    errD_fake = criterion(output, label)
    errD_fake.backward()
    optimizerD.step()
    # ---- "decoder" phase (:127-164)
    netEG.zero_grad()
    label.fill_(real_label)
    output, sim_real = netD(data)
    recon, mu, logvar = _vae_forward(netEG, data, eps_dec)
    output_fake, _ = netD(fake)
    output_recon, sim_recon = netD(recon)
    errG_fake = criterion(output_fake, label)
    errG_recon = criterion(output_recon, label)
    errG_fake.backward(retain_graph=True)
    errG_recon.backward(retain_graph=True)
    sim_loss = 0.5 * F.mse_loss(sim_recon, sim_real, reduction="sum")
    sim_loss.backward(retain_graph=True)
    loss_dec = F.mse_loss(recon, data, reduction="sum")
    loss_dec.backward()
    optimizerEG.step()
    # ---- "encoder" phase (:167-193)
    netEG.zero_grad()
    recon, mu, logvar = _vae_forward(netEG, data, eps_enc)
    kld = beta * kld_sum(mu, logvar)
    kld.backward(retain_graph=True)
    loss_enc = F.mse_loss(recon, data, reduction="sum")
    loss_enc.backward()
    optimizerEG.step()
    return {"errD_real": errD_real.item(), "errD_fake": errD_fake.item(), "D_x": D_x,
            "errG_fake": errG_fake.item(), "errG_recon": errG_recon.item(), "sim": sim_loss.item(),
            "recon_dec": loss_dec.item(), "kld": kld.item(), "recon_enc": loss_enc.item()}


def make_opt(n_hidden=128, n_z=(256, 8, 8), input_channels=3):
    """The three fields the models read from EnvSetter's namespace (utils/envsetter.py:34,41-42)."""
    from types import SimpleNamespace

    return SimpleNamespace(input_channels=input_channels, n_hidden=n_hidden, n_z=list(n_z))


def synthetic_batch(batch, seed, device="cpu"):
    """CelebA after Normalize(.5,.5) lies in [-1,1] (dataloader/dataset.py:12,38-43): i.i.d. U[-1,1] stand-in."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 3, 64, 64, generator=g) * 2 - 1).to(device)


class PerShard:
    """Data-parallel restatement for the N-rank parity check: every call of the wrapped
    module runs shard by shard over `world` equal slices of the batch -- so BatchNorm normalises each shard with
    its own statistics, as each replica does under the reference's nn.DataParallel (new_betavaegan.py:40-42) and
    as each rank does in the one-process-per-GPU scheme -- and the outputs are concatenated in rank order, so a
    loss taken over the concatenated outputs back-propagates the SUM of the per-shard gradients.  Only shard 0's
    forward updates the persistent BatchNorm buffers (DataParallel keeps device 0's; rank 0 is what the check
    compares).  Pass instances of this class to the *_step functions above in place of the modules."""

    def __init__(self, module, world):
        self.module, self.world = module, world

    def _run(self, fn, x):
        outs = []
        for s, xs in enumerate(x.chunk(self.world)):
            saved = []
            if s > 0:  # shards > 0 update throw-away copies of the BatchNorm buffers
                for m in self.module.modules():
                    for k, b in list(m._buffers.items()):
                        if b is not None:
                            saved.append((m, k, b))
                            m._buffers[k] = b.clone()
            o = fn(xs)
            for m, k, b in saved:
                m._buffers[k] = b
            outs.append(o if isinstance(o, tuple) else (o,))
        cat = tuple(torch.cat([o[i] for o in outs]) for i in range(len(outs[0])))
        return cat if len(cat) > 1 else cat[0]

    def __call__(self, x):
        return self._run(self.module, x)

    def encode(self, x):
        return self._run(self.module.encode, x)

    def decode(self, z):
        return self._run(self.module.decode, z)

    def zero_grad(self):
        self.module.zero_grad()

    def parameters(self):
        return self.module.parameters()

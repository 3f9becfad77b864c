"""Drop-in for the reference's `models/model.py` (CelebA family): same class names, constructor signature,
attribute names, state_dict keys and return values — but `forward` runs the libdm_b200 CUDA kernels through
custom autograd.Functions instead of torch.nn's library ops.

    from model import *          # reference:  experiments/new_vae.py:13, new_gan.py:22, new_betavaegan.py:18
    -> put `dropin/` on sys.path (it re-exports this module as `model`).

The submodules (`features`, `x_to_mu`, ..., `deconv4`, `activation`, `convs`, `lth_features`,
`sigmoid_output`) are ordinary torch.nn layers used as PARAMETER HOLDERS, so `.apply(weights_init)`,
`.to()`, `.state_dict()`, `DataParallel(...).module`, optimizers and checkpoints behave exactly as in the
reference (models/model.py:282-571).  They are never called.

There is no CPU path: inputs must be CUDA tensors and libdm_b200.so must load, otherwise this raises.
"""
from __future__ import annotations

import torch
from torch import nn

from . import engine, ops

__all__ = ["weights_init", "VAE", "Discriminator_celeba", "Generator_celeba", "Encoder_celeba"]


def default_opt():
    """The three fields the models read from the reference's EnvSetter namespace (utils/envsetter.py:34,41-42)."""
    from types import SimpleNamespace

    return SimpleNamespace(input_channels=3, n_hidden=128, n_z=[256, 8, 8])

K, PAD = 5, 2


def weights_init(m):
    """models/model.py:8-14."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: disentangle_mlp_b200 has no CPU path; got a {t.device} tensor")


def _alloc_grads(P):
    """One zeroed flat fp32 buffer, one view per parameter (gradients are accumulated into them)."""
    names = list(P)
    sizes = [(P[n].numel() + 63) // 64 * 64 for n in names]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=P[names[0]].device)
    G, off = {}, 0
    for n, s in zip(names, sizes):
        G[n] = flat[off:off + P[n].numel()].view(P[n].shape)
        off += s
    return G


class _NetFn(torch.autograd.Function):
    """Shared plumbing: forward(ctx, owner, kind, x, *params)."""

    @staticmethod
    def forward(ctx, owner, kind, x, *params):
        names = owner._param_names(kind)
        P = dict(zip(names, params))
        B = dict(owner.named_buffers())
        cache = owner._operand_cache
        ctx.owner, ctx.kind, ctx.names, ctx.P = owner, kind, names, P
        # backward re-derives the bf16 operands from the CURRENT parameters (they are not saved on ctx): remember
        # their versions so that an in-place update between forward and backward raises, as stock autograd would
        ctx.versions = [(p.data_ptr(), p._version) for p in params]
        x = x.detach()
        if kind == "disc":
            prob, feat, S = engine.discriminator_forward(x.float().contiguous(), P, B, cache, owner.training)
            ctx.S = S
            return prob, feat
        if kind == "enc":
            mu, logvar, S = engine.encoder_forward(x.float().contiguous(), P, B, cache, owner.training)
            ctx.S = S
            return mu, logvar
        recon, S = engine.decoder_forward(x.float().contiguous(), P, B, cache, owner.training)
        ctx.S = S
        return recon

    @staticmethod
    def backward(ctx, *grads):
        P, names, cache = ctx.P, ctx.names, ctx.owner._operand_cache
        if ctx.versions != [(P[n].data_ptr(), P[n]._version) for n in names]:
            raise RuntimeError(f"{type(ctx.owner).__name__}: a parameter was modified in place between forward and "
                               "backward (optimizer.step() before a retained-graph backward?); the kernels would "
                               "back-propagate through the updated weights")
        need_w = any(ctx.needs_input_grad[3:])
        G = _alloc_grads(P) if need_w else None
        need_dx = ctx.needs_input_grad[2]
        grads = [None if g is None else g.contiguous().float() for g in grads]
        if ctx.kind == "disc":
            dx = engine.discriminator_backward(ctx.S, grads[0], grads[1], P, G, cache, need_dx, need_w)
        elif ctx.kind == "enc":
            dx = engine.encoder_backward(ctx.S, grads[0], grads[1], P, G, cache, need_w)
        else:
            g = grads[0] if grads[0] is not None else torch.zeros_like(ctx.S.recon)
            dx = engine.decoder_backward(ctx.S, g, P, G, cache, need_dx, need_w)
        pg = [G[n] if (G is not None and ctx.needs_input_grad[3 + i]) else None for i, n in enumerate(names)]
        return (None, None, dx, *pg)


class _ReparamFn(torch.autograd.Function):
    """z = mu + eps * exp(0.5 * logvar)  (models/model.py:532-535)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        mu, logvar, eps = mu.contiguous(), logvar.contiguous(), eps.contiguous()
        z, _ = ops.reparam_forward(mu, logvar, eps)
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        _, _, dmu, dlogvar = ops.reparam_backward(dz.contiguous(), logvar, eps)
        return dmu, dlogvar, None


class _KernelBacked(nn.Module):
    """Mixin state shared by the drop-in modules."""

    _ENC_PREFIXES = ("features.", "x_to_mu.", "x_to_logvar.")

    def _init_backend(self):
        self._operand_cache = engine.OperandCache()

    def _param_names(self, kind):
        names = [n for n, _ in self.named_parameters()]
        if kind == "enc":
            return [n for n in names if n.startswith(self._ENC_PREFIXES)]
        if kind == "dec":
            return [n for n in names if not n.startswith(self._ENC_PREFIXES)]
        return names

    def _run(self, kind, x):
        _require_cuda(x, type(self).__name__)
        names = self._param_names(kind)
        P = dict(self.named_parameters())
        return _NetFn.apply(self, kind, x, *[P[n] for n in names])

    def apply(self, fn):  # .apply(weights_init) writes through `.data`: version counters do not move
        out = super().apply(fn)
        if hasattr(self, "_operand_cache"):
            self._operand_cache.invalidate()
        return out

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.float(): parameters move, derived operands are stale
        out = super()._apply(fn, *a, **k)
        if hasattr(self, "_operand_cache"):
            self._operand_cache.invalidate()
        return out


def _conv_stack(channels, strides, act):
    layers = []
    for cin, cout, s in zip(channels[:-1], channels[1:], strides):
        layers += [nn.Conv2d(cin, cout, K, stride=s, padding=PAD), nn.BatchNorm2d(cout), act()]
    return nn.Sequential(*layers)


def _latent_head(n_in, n_hidden):
    return nn.Sequential(nn.Linear(n_in, 2048), nn.BatchNorm1d(2048), nn.ReLU(), nn.Linear(2048, n_hidden))


def _check_supported(opt, who):
    nz = list(opt.n_z)
    if opt.input_channels != 3 or opt.n_hidden != 128 or nz != [256, 8, 8]:
        raise NotImplementedError(
            f"{who}: the B200 kernels are specialised to the reference's CelebA configuration "
            f"(input_channels=3, n_hidden=128, n_z=[256,8,8]; utils/envsetter.py:34,41-42); got "
            f"input_channels={opt.input_channels}, n_hidden={opt.n_hidden}, n_z={nz}")


class _DecoderLayers:
    def _build_decoder(self, opt):
        nz = list(opt.n_z)
        dim = nz[0] * nz[1] * nz[2]
        self.preprocess = nn.Sequential(nn.Linear(opt.n_hidden, dim), nn.BatchNorm1d(dim), nn.ReLU())
        self.deconv1 = nn.ConvTranspose2d(nz[0], 256, K, stride=2, padding=PAD)
        self.act1 = nn.Sequential(nn.BatchNorm2d(256), nn.ReLU())
        self.deconv2 = nn.ConvTranspose2d(256, 128, K, stride=2, padding=PAD)
        self.act2 = nn.Sequential(nn.BatchNorm2d(128), nn.ReLU())
        self.deconv3 = nn.ConvTranspose2d(128, 32, K, stride=2, padding=PAD)
        self.act3 = nn.Sequential(nn.BatchNorm2d(32), nn.ReLU())
        self.deconv4 = nn.ConvTranspose2d(32, 3, K, stride=1, padding=PAD)
        self.activation = nn.Tanh()


class Generator_celeba(_KernelBacked, _DecoderLayers):
    """models/model.py:331-378."""

    def __init__(self, opt):
        super().__init__()
        _check_supported(opt, "Generator_celeba")
        self.input_size = opt.n_hidden
        self.representation_size = opt.n_z
        self._build_decoder(opt)
        self._init_backend()

    def forward(self, code):
        return self._run("dec", code)


class Discriminator_celeba(_KernelBacked):
    """models/model.py:381-416 — returns (probability [B], Dis_l features [B,2048])."""

    def __init__(self, opt):
        super().__init__()
        _check_supported(opt, "Discriminator_celeba")
        self.representation_size = opt.n_z
        dim = opt.n_z[0] * opt.n_z[1] * opt.n_z[2]
        self.convs = _conv_stack([opt.input_channels, 32, 128, 256, 256], [1, 2, 2, 2], lambda: nn.LeakyReLU(0.2))
        self.lth_features = nn.Sequential(nn.Linear(dim, 2048), nn.LeakyReLU(0.2))
        self.sigmoid_output = nn.Sequential(nn.Linear(2048, 1), nn.Sigmoid())
        self._init_backend()

    def forward(self, x):
        prob, feat = self._run("disc", x)
        return prob.squeeze(), feat.squeeze()


class Encoder_celeba(_KernelBacked):
    """models/model.py:282-328 — returns (z, per-sample KL)."""

    def __init__(self, opt, representation_size=64):
        super().__init__()
        _check_supported(opt, "Encoder_celeba")
        if representation_size != 64:
            raise NotImplementedError("Encoder_celeba: representation_size must be 64")
        self.input_channels = opt.input_channels
        self.n_hidden = opt.n_hidden
        r = representation_size
        self.features = _conv_stack([opt.input_channels, r, 2 * r, 4 * r], [2, 2, 2], nn.ReLU)
        self.x_to_mu = _latent_head(4 * r * 8 * 8, opt.n_hidden)
        self.x_to_logvar = _latent_head(4 * r * 8 * 8, opt.n_hidden)
        self._init_backend()

    def forward(self, x):
        mu, logvar = self._run("enc", x)
        eps = torch.randn(mu.size()).to(mu.device)  # model.py:319 draws on the host, then moves
        z = _ReparamFn.apply(mu, logvar, eps)
        kld = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), 1)
        return z, kld


class VAE(_KernelBacked, _DecoderLayers):
    """models/model.py:419-571 — encoder + decoder in one module."""

    def __init__(self, opt, representation_size=64):
        super().__init__()
        _check_supported(opt, "VAE")
        if representation_size != 64:
            raise NotImplementedError("VAE: representation_size must be 64")
        self.input_channels = opt.input_channels
        self.n_hidden = opt.n_hidden
        r = representation_size
        self.features = _conv_stack([opt.input_channels, r, 2 * r, 4 * r], [2, 2, 2], nn.ReLU)
        self.x_to_mu = _latent_head(4 * r * 8 * 8, opt.n_hidden)
        self.x_to_logvar = _latent_head(4 * r * 8 * 8, opt.n_hidden)
        self.input_size = opt.n_hidden
        self.representation_size2 = opt.n_z
        self._build_decoder(opt)
        self._init_backend()

    def encode(self, x):
        return self._run("enc", x)

    def reparameterize(self, mu, logvar):
        eps = torch.randn_like(mu)  # model.py:534 (randn_like(std): same shape/device/dtype as mu)
        return _ReparamFn.apply(mu, logvar, eps)

    def decode(self, code):
        return self._run("dec", code)

    def forward(self, x):
        mu, logvar = self.encode(x)
        z = self.reparameterize(mu, logvar)
        return self.decode(z), mu, logvar

"""Sampling / reconstruction glue of the reference: utils/utils.py:6-32, called once per epoch by
experiments/new_betavaegan.py:233,257-265, new_gan.py:178-183, new_vae.py:94-110 and by
utils/generate_samples_recons.py:36-56.  Same function names, argument order and file names; `fn` is the caller's
closure over the kernel-backed modules (`lambda z: netEG.module.decode(z).cpu()`, `lambda x: netEG(x.to(dev))[0]`),
which run their forward-only CUDA path under torch.no_grad().  BatchNorm stays in training mode there, exactly as in
the reference (its scripts never call .eval())."""
from __future__ import annotations

import torch

__all__ = ["gen_fid_reconstructions", "gen_reconstructions", "generate_fid_samples", "generate_samples"]


def _save_image(t, path, **kw):
    from torchvision.utils import save_image  # torchvision is the reference's own dependency for this (utils.py:2)

    if str(path).endswith(".pdf"):  # Pillow's PDF writer needs its JPEG plugin registered first
        from PIL import Image

        Image.init()
    save_image(t, path, **kw)


def gen_fid_reconstructions(fn, dl, epoch, results_path, ext="pdf"):
    """utils/utils.py:6-11 — one file per reconstructed image of the loader's first batch."""
    with torch.no_grad():
        orig_imgs, _ = next(iter(dl))
        batch = fn(orig_imgs).cpu()
        for i, x in enumerate(batch):
            _save_image(x.cpu(), results_path + f"/recon_{i}_{str(epoch)}.{ext}", normalize=True)
    return batch


def gen_reconstructions(fn, dl, epoch, results_path, nrow=8, path_for_originals="", ext="pdf"):
    """utils/utils.py:13-19 — one grid of reconstructions (and optionally the originals)."""
    with torch.no_grad():
        orig_imgs, _ = next(iter(dl))
        batch = fn(orig_imgs).cpu()
        _save_image(batch.cpu(), results_path + f"/recon_{str(epoch)}.{ext}", nrow=nrow, normalize=True)
        if path_for_originals:
            _save_image(orig_imgs.cpu(), path_for_originals + f"/original_{str(epoch)}.{ext}", nrow=nrow, normalize=True)
    return batch


def generate_fid_samples(fn, epoch, n_samples, n_hidden, results_path, device="cpu", ext="pdf"):
    """utils/utils.py:21-26 — noise is drawn on the host, then moved (same torch RNG stream as the reference)."""
    with torch.no_grad():
        sample = torch.randn(n_samples, n_hidden).to(device)
        sample = fn(sample).cpu()
        for i, x in enumerate(sample):
            _save_image(x.cpu(), results_path + f"/sample_{i}_{str(epoch)}.{ext}", normalize=True)
    return sample


def generate_samples(fn, epoch, n_samples, n_hidden, results_path, nrow=8, device="cpu", ext="pdf"):
    """utils/utils.py:28-32."""
    with torch.no_grad():
        sample = torch.randn(n_samples, n_hidden).to(device)
        sample = fn(sample).cpu()
        _save_image(sample.cpu(), results_path + f"/sample_{str(epoch)}.{ext}", nrow=nrow, normalize=True)
    return sample

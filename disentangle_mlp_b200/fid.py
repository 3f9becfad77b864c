"""FID scoring on the GPU — the reference's only quality metric (scoring/fid.py, called once per epoch
by experiments/new_*.py with --calc_fid true).

  frechet_distance        scoring/fid.py:109-160.  The reference takes scipy's sqrtm of the NON-symmetric 2048x2048 product
                          C1 C2 on one CPU core (~30 s).  Here, in fp64 on the GPU: C1 C2 is similar to the symmetric PSD
                          matrix  S = C1^(1/2) C2 C1^(1/2)  (same eigenvalues), so  Tr sqrt(C1 C2) = sum sqrt(eig(S)) --
                          two symmetric eigendecompositions (torch.linalg.eigh) and two GEMMs, no complex arithmetic and
                          no "imaginary component" failure mode.
  activation_statistics   scoring/fid.py:165-184: mean and unbiased covariance of the pool_3 activations (fp64).
  InceptionPool3          the 2048-d pool_3 feature extractor (scoring/inception.py:16-190 wraps torchvision's
                          Inception-v3).  The architecture is torchvision's; the FID weights
                          (pt_inception-2015-12-05, scoring/inception.py:13) cannot be downloaded in this sandbox, so a
                          local state_dict path is required for meaningful scores -- without it the extractor is randomly
                          initialised and says so.
  get_fid                 scoring/fid.py:303-323: images of a directory (or a uint8 .npy stack) against precomputed
                          statistics (.npz with mu, sigma).

This is glue around library linear algebra (cuSOLVER / cuBLAS through torch), run once per epoch -- not part of the
training step's hot path, and deliberately not hand-written kernels."""
from __future__ import annotations

import os

import numpy as np
import torch

__all__ = ["frechet_distance", "activation_statistics", "InceptionPool3", "get_activations", "get_fid"]


def _sym_sqrt(c):
    w, v = torch.linalg.eigh(c)
    return (v * w.clamp_min(0).sqrt()) @ v.T


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6, device=None):
    """d^2 = ||mu1 - mu2||^2 + Tr(C1 + C2 - 2 sqrt(C1 C2)), fp64 on `device` (default: cuda if available)."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=device)  # noqa: E731
    mu1, mu2, c1, c2 = t(mu1).flatten(), t(mu2).flatten(), t(sigma1), t(sigma2)
    assert mu1.shape == mu2.shape and c1.shape == c2.shape and c1.shape == (mu1.numel(), mu1.numel())
    c1, c2 = 0.5 * (c1 + c1.T), 0.5 * (c2 + c2.T)
    r = _sym_sqrt(c1)
    ev = torch.linalg.eigvalsh(r @ c2 @ r)
    if not bool(torch.isfinite(ev).all()):  # (the reference's fallback: regularise both covariances, fid.py:144-148)
        eye = torch.eye(c1.shape[0], dtype=torch.float64, device=device) * eps
        r = _sym_sqrt(c1 + eye)
        ev = torch.linalg.eigvalsh(r @ (c2 + eye) @ r)
    tr_covmean = ev.clamp_min(0).sqrt().sum()
    diff = mu1 - mu2
    return float(diff.dot(diff) + torch.trace(c1) + torch.trace(c2) - 2 * tr_covmean)


def activation_statistics(act, device=None):
    """(mu, sigma) of activations [n, d]: np.mean(axis=0), np.cov(rowvar=False) -- fp64, returned as numpy arrays."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    a = torch.as_tensor(np.asarray(act) if not torch.is_tensor(act) else act).to(device=device, dtype=torch.float64)
    mu = a.mean(dim=0)
    d = a - mu
    sigma = d.T @ d / (a.shape[0] - 1)
    return mu.cpu().numpy(), sigma.cpu().numpy()


class InceptionPool3(torch.nn.Module):
    """Inception-v3 up to the final average pool (2048 features), inputs in [0, 1], resized to 299x299 and scaled to
    [-1, 1] as scoring/inception.py:129-160 does."""

    def __init__(self, weights_path=None):
        super().__init__()
        from torchvision import models

        net = models.inception_v3(weights=None, aux_logits=True, transform_input=False, init_weights=False)
        self.pretrained = False
        if weights_path:
            sd = torch.load(weights_path, map_location="cpu")
            missing, unexpected = net.load_state_dict(sd, strict=False)
            self.pretrained = True
        net.fc = torch.nn.Identity()
        net.eval()
        self.net = net
        for p in self.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def forward(self, x):
        x = torch.nn.functional.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
        return self.net(2 * x - 1)


@torch.no_grad()
def get_activations(images_u8, model, batch_size=50, device="cuda"):
    """images_u8: uint8 [n, h, w, 3] (values 0..255, as scoring/fid.py:83-103 expects) -> [n, 2048] float64"""
    out = []
    model = model.to(device)
    for i in range(0, len(images_u8), batch_size):
        b = torch.as_tensor(np.asarray(images_u8[i:i + batch_size])).to(device).permute(0, 3, 1, 2).float() / 255.0
        out.append(model(b).double().cpu())
    return torch.cat(out).numpy()


def _load_images(path):
    if path.endswith(".npy"):
        return np.load(path)
    from PIL import Image

    files = sorted(f for f in os.listdir(path) if f.lower().endswith((".png", ".jpg", ".jpeg", ".bmp")))
    return np.stack([np.asarray(Image.open(os.path.join(path, f)).convert("RGB")) for f in files])


def get_fid(path_data, path_pretrained, inception="", lowprofile=False, device="cuda"):
    """scoring/fid.py:303-323: FID of the images under `path_data` against the statistics file `path_pretrained`
    (.npz with mu, sigma).  `inception`: path of a local Inception-v3 state_dict (required for meaningful scores)."""
    model = InceptionPool3(inception or None)
    if not model.pretrained:
        import warnings

        warnings.warn("get_fid: no Inception weights given (they cannot be downloaded here): features come from a "
                      "randomly initialised network and the score is NOT comparable with published FID values")
    act = get_activations(_load_images(path_data), model, device=device)
    mu, sigma = activation_statistics(act, device=device)
    ref = np.load(path_pretrained)
    return frechet_distance(mu, sigma, ref["mu"], ref["sigma"], device=device)

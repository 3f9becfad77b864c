"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU restatement (numpy + scipy, fp64) of the arithmetic of the reference's FID score, scoring/fid.py:
  frechet_distance        scoring/fid.py:109-160 (calculate_frechet_distance, "stable version by Dougal J. Sutherland")
  activation_statistics   scoring/fid.py:165-184 (mean and np.cov(rowvar=False) of the pool_3 activations)
The reference module itself cannot be imported here (it imports TensorFlow at the top, scoring/fid.py:20-30, and
downloads the Inception graph); the two functions below follow it statement by statement.  Pinned by closed-form
known answers in tests/test_fid.py (identical Gaussians, commuting covariances)."""
from __future__ import annotations

import warnings

import numpy as np
from scipy import linalg


def _sqrtm(a):
    """linalg.sqrtm(a, disp=False)[0] in the reference; the `disp` argument is gone from the scipy of this image
    (1.18), where sqrtm(a) returns the matrix alone."""
    try:
        return linalg.sqrtm(a, disp=False)[0]
    except TypeError:
        return linalg.sqrtm(a)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """d^2 = ||mu1 - mu2||^2 + Tr(C1 + C2 - 2 sqrt(C1 C2))   (scoring/fid.py:109-160)"""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = _sqrtm(sigma1.dot(sigma2))  # product might be almost singular
    if not np.isfinite(covmean).all():
        warnings.warn("fid calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):  # numerical error might give a slight imaginary component
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def activation_statistics(act):
    """scoring/fid.py:181-184"""
    act = np.asarray(act, dtype=np.float64)
    return np.mean(act, axis=0), np.cov(act, rowvar=False)

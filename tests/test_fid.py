"""FID arithmetic: the oracle's restatement of scoring/fid.py against closed-form known answers (CPU),
and the GPU implementation (symmetric-eigendecomposition form, fp64) against the oracle."""
import numpy as np
import pytest
import torch


def _psd(d, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((d, 2 * d))
    return scale * (a @ a.T) / (2 * d)


def test_oracle_frechet_known_answers():
    from oracle import fid as ofid

    d = 48
    c = _psd(d, 0)
    mu = np.linspace(-1, 1, d)
    assert abs(ofid.frechet_distance(mu, c, mu, c)) < 1e-6  # identical Gaussians
    assert abs(ofid.frechet_distance(mu, c, mu + 0.5, c) - 0.25 * d) < 1e-6  # pure mean shift
    # commuting (diagonal) covariances: sum (sqrt(a) - sqrt(b))^2
    a, b = np.linspace(0.5, 2.0, d), np.linspace(2.0, 0.1, d)
    want = float(((np.sqrt(a) - np.sqrt(b)) ** 2).sum())
    assert abs(ofid.frechet_distance(mu, np.diag(a), mu, np.diag(b)) - want) < 1e-8
    act = np.random.default_rng(1).standard_normal((500, d)) @ np.linalg.cholesky(c).T + mu
    m, s = ofid.activation_statistics(act)
    assert np.allclose(m, act.mean(0)) and np.allclose(s, np.cov(act.T))


def test_frechet_cpu_matches_oracle():
    from disentangle_mlp_b200 import fid
    from oracle import fid as ofid

    for d, seed in ((32, 3), (128, 4)):
        c1, c2 = _psd(d, seed), _psd(d, seed + 10, 1.7)
        m1, m2 = np.random.default_rng(seed).standard_normal(d), np.random.default_rng(seed + 1).standard_normal(d)
        want = ofid.frechet_distance(m1, c1, m2, c2)
        got = fid.frechet_distance(m1, c1, m2, c2, device="cpu")
        assert abs(got - want) <= 1e-8 * max(1.0, abs(want)), (d, got, want)


@pytest.mark.gpu
def test_fid_on_gpu_matches_oracle(tmp_path):
    from disentangle_mlp_b200 import fid
    from oracle import fid as ofid

    d = 512
    rng = np.random.default_rng(5)
    c = _psd(d, 6)
    act1 = rng.standard_normal((1500, d)) @ np.linalg.cholesky(c).T
    act2 = 1.2 * rng.standard_normal((1200, d)) @ np.linalg.cholesky(_psd(d, 7)).T + 0.1
    m1, s1 = fid.activation_statistics(act1)
    m2, s2 = fid.activation_statistics(act2)
    om1, os1 = ofid.activation_statistics(act1)
    assert np.allclose(m1, om1, atol=1e-10) and np.allclose(s1, os1, atol=1e-10)
    want = ofid.frechet_distance(m1, s1, m2, s2)
    got = fid.frechet_distance(m1, s1, m2, s2, device="cuda")
    assert abs(got - want) <= 1e-7 * abs(want), (got, want)
    # rank-deficient covariance (fewer samples than dimensions): the case the reference patches with eps, fid.py:144-148
    m3, s3 = fid.activation_statistics(act1[:200])
    got = fid.frechet_distance(m3, s3, m2, s2, device="cuda")
    want = ofid.frechet_distance(m3, s3, m2, s2)
    assert abs(got - want) <= 1e-4 * abs(want), (got, want)
    # end-to-end glue over the feature extractor (random weights offline: plumbing only)
    imgs = (rng.random((20, 64, 64, 3)) * 255).astype(np.uint8)
    np.save(tmp_path / "imgs.npy", imgs)
    model = fid.InceptionPool3()
    act = fid.get_activations(imgs, model, batch_size=8)
    assert act.shape == (20, 2048) and np.isfinite(act).all()
    mu, sigma = fid.activation_statistics(act)
    np.savez(tmp_path / "stats.npz", mu=mu, sigma=sigma)
    with pytest.warns(UserWarning):
        assert np.isfinite(fid.get_fid(str(tmp_path / "imgs.npy"), str(tmp_path / "stats.npz")))

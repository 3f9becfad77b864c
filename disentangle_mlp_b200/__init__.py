"""disentangle_mlp_b200 — B200-native kernels behind the VAE / GAN / beta-VAE-GAN training step of
RicoFio/disentangle_mlp (drop-in `model.py` classes + fused training steps)."""

__version__ = "0.1.0"

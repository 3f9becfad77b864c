"""VAE training on the B200 kernels — counterpart of the reference's experiments/new_vae.py (checkpoint keys
:88-91; loop body = disentangle_mlp_b200.trainer.VAETrainer.step)."""
import os

import torch

from _common import Loader, parse, setup_dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200.trainer import VAETrainer


def main():
    opt = parse("vae")
    world, rank, dev = setup_dist()
    torch.manual_seed(opt.seed)
    model = dm.VAE(opt).to(dev)
    model.apply(dm.weights_init)
    T = VAETrainer(model, lr=opt.lr)
    torch.manual_seed(opt.seed + 7919 * (rank + 1))  # same initial weights on every rank, different noise / eps
    loader = Loader(opt, world, rank, dev)
    for epoch in range(opt.epochs):
        total = None
        for i, data in enumerate(loader):
            m = T.step(data)
            total = m["loss"] if total is None else total + m["loss"]
            if rank == 0 and i % opt.log_interval == 0:
                print(f"Train Epoch: {epoch} [{i}/{len(loader)}]\tLoss: {float(m['loss']) / data.shape[0]:.6f}", flush=True)
        if rank == 0:
            print(f"====> Epoch: {epoch} Average loss: {float(total) / loader.dataset_len * world:.4f}")
            if opt.model_path:
                os.makedirs(opt.model_path, exist_ok=True)
                torch.save({"epoch": epoch + 1, "VAE_model": model.state_dict(), "optimizer": T.fp.optimizer_state_dict()},
                           os.path.join(opt.model_path, f"model_{epoch + 1}.tar"))


if __name__ == "__main__":
    main()

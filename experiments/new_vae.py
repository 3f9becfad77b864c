"""VAE training on the B200 kernels — counterpart of the reference's experiments/new_vae.py (checkpoint keys
:88-91; loop body = disentangle_mlp_b200.trainer.VAETrainer.step)."""
import torch

from _common import Loader, load_checkpoint, parse, save_checkpoint, setup_dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200.trainer import VAETrainer


def main():
    opt = parse("vae")
    world, rank, dev = setup_dist()
    torch.manual_seed(opt.seed)
    model = dm.VAE(opt).to(dev)
    model.apply(dm.weights_init)
    T = VAETrainer(model, lr=opt.lr)
    start = 0
    if opt.load_path:  # new_vae.py:72-76
        start = load_checkpoint("vae", opt.load_path, (model,), (T.fp,), dev)
    torch.manual_seed(opt.seed + 7919 * (rank + 1))  # same initial weights on every rank, different noise / eps
    loader = Loader(opt, world, rank, dev)
    for epoch in range(start, opt.epochs):
        total = None
        for i, data in enumerate(loader):
            m = T.step(data)
            total = m["loss"] if total is None else total + m["loss"]
            if rank == 0 and i % opt.log_interval == 0:
                print(f"Train Epoch: {epoch} [{i}/{len(loader)}]\tLoss: {float(m['loss']) / data.shape[0]:.6f}", flush=True)
        T.sync(masters=True)  # (data parallel: complete the sharded optimizer state before it is read)
        if rank == 0:
            print(f"====> Epoch: {epoch} Average loss: {float(total) / loader.dataset_len * world:.4f}")
            if opt.model_path:  # new_vae.py:88-91
                save_checkpoint("vae", opt.model_path, epoch + 1, (model,), (T.fp,))


if __name__ == "__main__":
    main()

"""beta-VAE-GAN training on the B200 kernels — counterpart of the reference's experiments/new_betavaegan.py.
Same models, losses, update order and checkpoint keys (:222-228); the loop body is
disentangle_mlp_b200.trainer.BetaVAEGANTrainer.step.  lr is hard-coded to 1e-3 as in the reference (:49-50).

    python experiments/new_betavaegan.py --name run --beta 25 --batch_size_train 64
    torchrun --nproc-per-node 8 experiments/new_betavaegan.py --name run --beta 25 --batch_size_train 512
"""
import numpy as np
import torch

from _common import Loader, load_checkpoint, parse, save_checkpoint, setup_dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200.trainer import BetaVAEGANTrainer


def main():
    opt = parse("vaegan")
    world, rank, dev = setup_dist()
    torch.manual_seed(opt.seed)
    np.random.seed(opt.seed)  # the label stream must be identical on every rank
    netEG, netD = dm.VAE(opt).to(dev), dm.Discriminator_celeba(opt).to(dev)
    netEG.apply(dm.weights_init)
    netD.apply(dm.weights_init)
    T = BetaVAEGANTrainer(netEG, netD, beta=opt.beta, lr=1e-3)
    start = 0
    if opt.load_path:  # new_betavaegan.py:203-209
        start = load_checkpoint("betavaegan", opt.load_path, (netEG, netD), (T.feg, T.fd), dev)
    torch.manual_seed(opt.seed + 7919 * (rank + 1))  # same initial weights on every rank, different noise / eps
    loader = Loader(opt, world, rank, dev)
    for epoch in range(start, opt.epochs):
        sums = None
        for i, data in enumerate(loader):
            m = T.step(data)
            vals = torch.stack([m["recon_enc"], m["recon_dec"], m["D_x"]])
            sums = vals if sums is None else sums + vals  # accumulated on the device: no per-step host sync
            if rank == 0 and i % opt.log_interval == 0:
                print(f"epoch {epoch} step {i}: " + " ".join(f"{k}={float(v):.4f}" for k, v in m.items()), flush=True)
        enc, dec, dx = (float(v) / loader.dataset_len * world for v in sums)
        T.sync(masters=True)  # (data parallel: complete the sharded optimizer state before it is read)
        if rank == 0:
            print(f"====> Epoch: {epoch} Avg Encoder Loss: {enc:.4f} Avg Decoder Loss: {dec:.4f} Dx: {dx:.4f}")
            if opt.model_path:  # new_betavaegan.py:222-228
                save_checkpoint("betavaegan", opt.model_path, epoch + 1, (netEG, netD), (T.feg, T.fd))


if __name__ == "__main__":
    main()

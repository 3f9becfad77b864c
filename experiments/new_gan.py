"""GAN training on the B200 kernels — counterpart of the reference's experiments/new_gan.py (checkpoint keys
:169-174; loop body = disentangle_mlp_b200.trainer.GANTrainer.step)."""
import numpy as np
import torch

from _common import Loader, load_checkpoint, parse, save_checkpoint, setup_dist

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200.trainer import GANTrainer


def main():
    opt = parse("gan")
    world, rank, dev = setup_dist()
    torch.manual_seed(opt.seed)
    np.random.seed(opt.seed)
    netG, netD = dm.Generator_celeba(opt).to(dev), dm.Discriminator_celeba(opt).to(dev)
    netG.apply(dm.weights_init)
    netD.apply(dm.weights_init)
    T = GANTrainer(netG, netD, lr=opt.lr)
    start = 0
    if opt.load_path:  # new_gan.py:143-151
        start = load_checkpoint("gan", opt.load_path, (netG, netD), (T.fg, T.fd), dev)
    torch.manual_seed(opt.seed + 7919 * (rank + 1))  # same initial weights on every rank, different noise / eps
    loader = Loader(opt, world, rank, dev)
    for epoch in range(start, opt.epochs):
        for i, data in enumerate(loader):
            m = T.step(data)
            if rank == 0 and i % opt.log_interval == 0:
                print("[%d/%d][%d/%d]\tLoss_D: %.4f\tLoss_G: %.4f\tD(x): %.4f\tD(G(z)): %.4f / %.4f" % (
                    epoch, opt.epochs, i, len(loader), float(m["errD"]), float(m["errG"]), float(m["D_x"]),
                    float(m["D_G_z1"]), float(m["D_G_z2"])), flush=True)
        T.sync(masters=True)  # (data parallel: complete the sharded optimizer state before it is read)
        if rank == 0 and opt.model_path:  # new_gan.py:169-174 (DataParallel state_dicts: "module." prefix)
            save_checkpoint("gan", opt.model_path, epoch + 1, (netG, netD), (T.fg, T.fd))


if __name__ == "__main__":
    main()

"""Shared glue of the three training scripts: flags (a subset of the reference's EnvSetter flags with the same
names and defaults, utils/envsetter.py:13-55), torchrun set-up, a synthetic CelebA-shaped loader, checkpoints."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse(desc):
    p = argparse.ArgumentParser(description=desc)
    p.add_argument("--name", required=True)
    p.add_argument("--seed", type=int, default=999)
    p.add_argument("--batch_size_train", type=int, default=256, help="GLOBAL batch (split across ranks)")
    p.add_argument("--n_z", type=int, nargs=3, default=[256, 8, 8])
    p.add_argument("--n_hidden", type=int, default=128)
    p.add_argument("--input_channels", type=int, default=3)
    p.add_argument("--lr", type=float, default=3e-3)
    p.add_argument("--beta", type=float, default=50)
    p.add_argument("--epochs", type=int, default=30)
    p.add_argument("--steps_per_epoch", type=int, default=100, help="synthetic data: steps per epoch")
    p.add_argument("--log_interval", type=int, default=10)
    p.add_argument("--model_path", default=None, help="directory for model_<epoch>.tar checkpoints")
    p.add_argument("--load_path", default=None)
    p.add_argument("--data_path", default=None, help="optional uint8 [N,64,64,3] .npy shard; default synthetic U[-1,1]")
    return p.parse_args()


def setup_dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: these scripts have no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, torch.device("cuda", local)


class Loader:
    """Per-rank shard of the training stream, already on the device.

    --data_path: a pre-decoded uint8 [N,64,64,3] .npy shard (memory-mapped).  Batches travel as uint8 NHWC -- a quarter
    of the fp32 bytes -- through a ring of pinned staging buffers with asynchronous H2D copies; the reference loader's
    ToTensor() + Normalize(.5,.5) (dataloader/dataset.py:37-43) runs on the GPU, fused into the training step's first
    kernel (dm_pad_image3).  At 12 k img/s/GPU the reference's ImageFolder + PIL + CPU-normalise pipeline
    (dataloader/dataset.py:45-50, 4 workers) is two orders of magnitude too slow.
    Default: synthetic CelebA-shaped fp32 batches in [-1, 1] (the dataset is not available offline)."""

    def __init__(self, opt, world, rank, dev, ring=4):
        self.b = opt.batch_size_train // world
        self.steps, self.dev = opt.steps_per_epoch, dev
        self.gen = torch.Generator().manual_seed(1234 + rank)
        self.data = None
        if opt.data_path:
            arr = np.load(opt.data_path, mmap_mode="r")
            assert arr.dtype == np.uint8 and arr.shape[1:] == (64, 64, 3), "expected a uint8 [N,64,64,3] shard"
            self.data = arr[rank::world]
            self.steps = len(self.data) // self.b
            self.ring = [(torch.empty((self.b, 64, 64, 3), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
                         for _ in range(ring)]
        self.dataset_len = self.steps * self.b * world

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            if self.data is None:
                x = torch.rand(self.b, 3, 64, 64, generator=self.gen) * 2 - 1
                yield x.pin_memory().to(self.dev, non_blocking=True)
            else:
                host, ev = self.ring[i % len(self.ring)]
                ev.synchronize()  # the copy that last read this staging buffer has completed
                host.copy_(torch.from_numpy(np.ascontiguousarray(self.data[i * self.b:(i + 1) * self.b])))
                x = host.to(self.dev, non_blocking=True)
                ev.record()
                yield x  # uint8 NHWC: trainer.step() normalises on the device


# ------------------------------------------------------------------------------------------ checkpoints
# The reference's `model_<epoch>.tar` dictionaries, key for key.  The reference wraps some networks
# in nn.DataParallel before calling .state_dict(), so THEIR keys carry a "module." prefix:
#   new_betavaegan.py:222-228  encoder_decoder_model = netEG.module.state_dict() (no prefix),
#                              discriminator_model = netD.state_dict()           ("module." prefix)
#   new_gan.py:169-174         netG / netD = DataParallel state_dicts            ("module." prefix on both)
#   new_vae.py:88-91           VAE_model = model.module.state_dict()             (no prefix)
# Written here with the same prefixes, so either side loads the other's files; on load a prefix is accepted or not.
CKPT_KEYS = {
    "betavaegan": (("encoder_decoder_model", False, "encoder_decoder_optimizer"),
                   ("discriminator_model", True, "discriminator_optimizer")),
    "gan": (("netG", True, "G_trainer"), ("netD", True, "D_trainer")),
    "vae": (("VAE_model", False, "optimizer"),),
}


def checkpoint_dict(kind, epoch, modules, flat_params):
    """modules / flat_params: in the order of CKPT_KEYS[kind] (e.g. (netEG, netD), (T.feg, T.fd))."""
    ck = {"epoch": epoch}
    for (mk, prefixed, ok), mod, fp in zip(CKPT_KEYS[kind], modules, flat_params):
        sd = mod.state_dict()
        ck[mk] = {("module." + k if prefixed else k): v for k, v in sd.items()}
        ck[ok] = fp.optimizer_state_dict()
    return ck


def save_checkpoint(kind, model_path, epoch, modules, flat_params):
    os.makedirs(model_path, exist_ok=True)
    path = os.path.join(model_path, f"model_{epoch}.tar")
    torch.save(checkpoint_dict(kind, epoch, modules, flat_params), path)
    return path


def load_checkpoint_dict(kind, ck, modules, flat_params):
    """Restore modules + fused optimizers from a reference-format dictionary; returns the epoch to resume from."""
    for (mk, _prefixed, ok), mod, fp in zip(CKPT_KEYS[kind], modules, flat_params):
        mod.load_state_dict({k.removeprefix("module."): v for k, v in ck[mk].items()})
        fp.load_optimizer_state_dict(ck[ok])
        fp.params_changed()
    return int(ck["epoch"])


def load_checkpoint(kind, path, modules, flat_params, dev):
    return load_checkpoint_dict(kind, torch.load(path, map_location=dev, weights_only=False), modules, flat_params)

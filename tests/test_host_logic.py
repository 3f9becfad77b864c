"""CPU: host-side logic of the drop-in modules and trainers (no kernels are launched)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from disentangle_mlp_b200 import model as dm
from oracle import nets, steps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_shapes_and_init_stream_match_reference():
    opt = steps.make_opt()
    for mine_cls, ref_cls in ((dm.VAE, nets.VAE), (dm.Discriminator_celeba, nets.Discriminator_celeba),
                              (dm.Generator_celeba, nets.Generator_celeba), (dm.Encoder_celeba, nets.Encoder_celeba)):
        torch.manual_seed(999)
        a = mine_cls(opt)
        a.apply(dm.weights_init)
        torch.manual_seed(999)
        b = ref_cls(opt)
        b.apply(nets.weights_init)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k


def test_attributes_the_training_scripts_touch_exist():
    m = dm.VAE(steps.make_opt())
    for name in ("features", "x_to_mu", "x_to_logvar", "preprocess", "deconv1", "act1", "deconv2", "act2", "deconv3",
                 "act3", "deconv4", "activation"):
        assert hasattr(m, name)
        getattr(m, name).requires_grad = False  # new_betavaegan.py:132-143 — a no-op attribute assignment
    assert all(p.requires_grad for p in m.parameters())
    assert len(m._param_names("enc")) == 24 and len(m._param_names("dec")) == 18
    assert sorted(m._param_names("enc") + m._param_names("dec")) == sorted(n for n, _ in m.named_parameters())


def test_unsupported_configuration_fails_loudly():
    from types import SimpleNamespace

    with pytest.raises(NotImplementedError):
        dm.VAE(SimpleNamespace(input_channels=1, n_hidden=128, n_z=[256, 8, 8]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dm.Discriminator_celeba(steps.make_opt())(torch.zeros(2, 3, 64, 64))


def test_dropin_shim_exports_reference_names():
    code = ("import sys; sys.path.insert(0, %r); from model import *; "
            "print(sorted(n for n in dir() if n in ('VAE','Discriminator_celeba','Generator_celeba','Encoder_celeba','weights_init')))"
            % os.path.join(ROOT, "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert "Discriminator_celeba" in out and "VAE" in out and "weights_init" in out


def test_label_stream_matches_reference_order():
    from disentangle_mlp_b200.trainer import _Base

    np.random.seed(5)
    a = [_Base.draw_labels() for _ in range(50)]
    np.random.seed(5)
    b = [steps.draw_labels() for _ in range(50)]
    assert a == b
    assert all(r in (0.1, 0.9) and f in (0.1, 0.9) for r, f in a)


def test_bench_reference_arm_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--workload", "vae"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0

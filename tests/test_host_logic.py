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


def _shell_flat_params(net, bf16_big=True):
    """A FlatParams WITHOUT device buffers (CPU tensors, no kernels): only what the all-reduce scheduling touches."""
    from types import SimpleNamespace

    import torch

    from disentangle_mlp_b200 import trainer as tr

    named = [(n, tuple(p.shape)) for n, p in net.named_parameters()]
    plan = tr.plan_layout(named, bf16_big)
    fp = object.__new__(tr.FlatParams)
    fp.names, fp.offsets, fp.total, fp._small_end = plan["names"], plan["offsets"], plan["total"], plan["small_end"]
    fp.off16, fp._late_ranges, fp._tail_done = plan["off16"], plan["late_ranges"], None
    fp.shard = False  # (the sharded reduce-scatter path of the big tensors is covered by tests/test_dist_cpu.py)
    fp.P = {n: SimpleNamespace(numel=lambda k=plan["numel"][n]: k) for n, _ in named}
    fp.grad = torch.zeros(plan["total"])
    fp.grad16 = torch.zeros(max(plan["total16"], 1), dtype=torch.bfloat16)
    return fp, plan


class _RecordingReducer:
    on = True

    def __init__(self):
        self.calls = []

    def allreduce_async(self, flat, lo=0, hi=None):
        self.calls.append((flat.dtype, lo, flat.numel() if hi is None else hi))

    def wait(self):
        pass


@pytest.mark.parametrize("bf16_big", [True, False])
def test_flat_layout_and_reduce_schedule_cover_every_gradient_once(bf16_big):
    """plan_layout + the trainer's all-reduce schedule (early big buckets, early decoder tail, phase-end rest): every
    gradient element is reduced exactly once per phase, the big tensors in bf16 when enabled; Adam segments and the
    zero ranges tile the flat buffer."""
    import numpy as np
    import torch

    from disentangle_mlp_b200 import trainer as tr
    from oracle import nets, steps

    opt = steps.make_opt()
    for net, early_names, tail in ((nets.VAE(opt), ("x_to_mu.0.weight", "x_to_logvar.0.weight"), "preprocess.0.weight"),
                                   (nets.Discriminator_celeba(opt), ("lth_features.0.weight",), None),
                                   (nets.Generator_celeba(opt), (), None)):
        fp, plan = _shell_flat_params(net, bf16_big)
        # layout: aligned, disjoint, big tensors last, small region contiguous
        spans = sorted((plan["offsets"][n], plan["offsets"][n] + plan["numel"][n]) for n in plan["names"])
        assert all(a % tr.ALIGN == 0 for a, _ in spans) and all(b <= c for (_, b), (c, _) in zip(spans, spans[1:]))
        assert all(plan["offsets"][n] >= plan["small_end"] for n in early_names)
        segs = plan["segments"]
        assert segs[0][0] == 0 and segs[-1][1] == plan["total"] and all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
        assert [s[2] is not None for s in segs].count(True) == (len(early_names) if bf16_big else 0)
        # one phase of the data-parallel schedule
        red = _RecordingReducer()
        for n in early_names:
            fp.reduce_early(red, n)
        if tail:
            fp.reduce_from(red, tail)
        fp.reduce_rest(red)
        cover32 = np.zeros(plan["total"], dtype=np.int8)
        cover16 = np.zeros(max(plan["total16"], 1), dtype=np.int8)
        for dt, lo, hi in red.calls:
            (cover16 if dt == torch.bfloat16 else cover32)[lo:hi] += 1
        for n in plan["names"]:
            o, k = plan["offsets"][n], plan["numel"][n]
            if n in plan["off16"]:
                o16 = plan["off16"][n]
                assert (cover16[o16:o16 + k] == 1).all() and (cover32[o:o + k] == 0).all(), n
            else:
                assert (cover32[o:o + k] == 1).all(), n
        assert len(red.calls) == len(early_names) + (1 if tail else 0) + 1
        assert fp._tail_done is None  # reset for the next phase


def test_the_two_head_weights_are_adjacent_in_the_flat_buffers():
    """engine.encoder_forward / encoder_backward run x_to_mu | x_to_logvar as ONE N = 4096 GEMM over the two 16384x2048
    weights (and one weight-gradient GEMM into their gradients): that needs them back to back, mu first, in the fp32
    flat buffer (= the bf16 shadow) and in the bf16 gradient buffer."""
    from disentangle_mlp_b200 import trainer as tr
    from oracle import nets, steps

    net = nets.VAE(steps.make_opt())
    named = [(n, tuple(p.shape)) for n, p in net.named_parameters()]
    for bf16_big in (True, False):
        plan = tr.plan_layout(named, bf16_big)
        a, b = "x_to_mu.0.weight", "x_to_logvar.0.weight"
        assert plan["offsets"][b] == plan["offsets"][a] + plan["numel"][a]
        if bf16_big:
            assert plan["off16"][b] == plan["off16"][a] + plan["numel"][a]

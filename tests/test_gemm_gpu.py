"""GPU: every mode of the tcgen05 tap-GEMM kernel through the C ABI against torch (fp32 math on the same bf16
inputs).  Tolerances: fp32 outputs 1e-4 relative-L2 (accumulation order only), bf16 outputs 4e-3 (one bf16
rounding of the result: 2^-9 per element)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import probe_gemm  # noqa: E402

pytestmark = pytest.mark.gpu

BF16_OUT = {"nt_bf16out", "nt_k128_n32", "nn_64x16384x2048"}


@pytest.mark.parametrize("name", list(probe_gemm.CASES))
def test_case(name):
    rel, _ = probe_gemm.CASES[name]()
    bf16_out = name in BF16_OUT or (name.startswith(("down_", "up_")) and "f32" not in name)
    assert rel < (4e-3 if bf16_out else 1e-4), (name, rel)


@pytest.mark.parametrize("env", [{"DM_CG2": "1"}, {"DM_CG2": "0"}])
@pytest.mark.parametrize("name", [n for n in probe_gemm.CASES if n.startswith(("down_", "up_", "wgrad_"))])
def test_cluster_modes(name, env, monkeypatch):
    """Convolution GEMMs as CTA pairs (DM_CG2=1, default: one tcgen05.mma.cta_group::2, M = 256, per pair; each CTA
    stages its 128 A rows and HALF of the B rows) and as single CTAs (DM_CG2=0)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rel, _ = probe_gemm.CASES[name]()
    bf16_out = name.startswith(("down_", "up_")) and "f32" not in name
    assert rel < (4e-3 if bf16_out else 1e-4), (name, env, rel)


def test_conv_up_c32_unmerged(monkeypatch):
    """The N = 32 four-phase form of the 128 -> 32 transposed convolution (default: phase-merged, N = 128)."""
    monkeypatch.setenv("DM_UP_MERGE", "0")
    rel, _ = probe_gemm.CASES["up_s2_c32"]()
    assert rel < 4e-3, rel


@pytest.mark.parametrize("batch,groups", [(4, 1), (64, 1), (96, 3)])
def test_conv_down_paired_c32(batch, groups):
    """Conv2d(32, 128, 5, stride 2, pad 2) (model.py:391) with two filter columns per k-block (dm_conv_down_paired,
    K = 64 rows of the parity view) against torch on the same bf16 operands, and against the one-column form."""
    import torch
    import torch.nn.functional as F

    from disentangle_mlp_b200 import ops

    torch.manual_seed(3)
    cs, cb, hs = 128, 32, 32
    x = torch.randn(batch, cb, 2 * hs, 2 * hs, device="cuda").bfloat16()
    w = (torch.randn(cs, cb, 5, 5, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(cs, device="cuda") * 0.1
    ref = F.conv2d(x.float(), w.float(), bias, stride=2, padding=2).permute(0, 2, 3, 1)
    g = ops.geom(batch, hs, hs, cs, cb, 2)
    wd, _, _ = ops.pack_conv_weights(w.float(), cs, cb)
    wp = ops.pack_down_pairs(wd, cs, cb)
    assert wp.shape == (15, cs, 64)
    xn = x.permute(0, 2, 3, 1).contiguous()
    y_pair = ops.conv_down(g, xn, wp, bias).float()
    y_one = ops.conv_down(g, xn, wd, bias).float()
    rel = float((y_pair - ref).norm() / ref.norm())
    assert rel < 4e-3, rel
    assert float((y_pair - y_one).norm() / y_one.norm()) < 4e-3
    if groups > 1 or batch >= 64:  # statistics in the epilogue ride along unchanged
        site = ops.BnSite(ops.bn_scratch(cs, groups, "cuda"), groups, batch // groups * hs * hs, cs,
                          torch.ones(cs, device="cuda"), torch.zeros(cs, device="cuda"), torch.zeros(cs, device="cuda"),
                          torch.ones(cs, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda"), 0.1, 1e-5)
        ops.conv_down(g, xn, wp, bias, bn=site)
        mean = ref.reshape(groups, -1, cs).mean(1)
        got = site.mean_invstd.view(groups, 2, cs)[:, 0]
        assert float((got - mean).abs().max()) < 2e-3 * float(ref.abs().max())


def test_full_size_layers_linearity_and_batch_independence():
    """Full BASELINE sizes (batch 128): conv(x1 + x2) == conv(x1) + conv(x2) on bf16-exact inputs, and the
    result for image i does not depend on the other images in the batch."""
    import torch

    from disentangle_mlp_b200 import ops

    torch.manual_seed(1)
    b = 128
    g = ops.geom(b, 16, 16, 256, 128, 2)
    x1 = torch.randint(-4, 5, (b, 32, 32, 128), device="cuda").float()
    x2 = torch.randint(-4, 5, (b, 32, 32, 128), device="cuda").float()
    # integers up to |8| and weights with few mantissa bits keep every product and partial sum exact in fp32
    wq = (torch.randint(-3, 4, (256, 128, 5, 5), device="cuda").float() / 8)
    wd, wu, _ = ops.pack_conv_weights(wq, 256, 128)
    y1 = ops.conv_down(g, x1.bfloat16(), wd).float()
    y2 = ops.conv_down(g, x2.bfloat16(), wd).float()
    y12 = ops.conv_down(g, (x1 + x2).bfloat16(), wd).float()
    assert torch.equal(y12, (y1 + y2).bfloat16().float()) or (y12 - (y1 + y2)).abs().max() <= 2 ** -7 * y12.abs().max()
    g1 = ops.geom(2, 16, 16, 256, 128, 2)
    ysub = ops.conv_down(g1, x1[5:7].bfloat16().contiguous(), wd).float()
    assert torch.equal(ysub, y1[5:7])
    # transposed convolution is the adjoint: <conv_down(x), s> == <x, conv_up(s)> (exact integers)
    s = torch.randint(-2, 3, (b, 16, 16, 256), device="cuda").float()
    up = ops.conv_up(g, s.bfloat16(), wu, out_f32=True)
    y = ops.conv_down(g, x1.bfloat16(), wd).double()
    lhs, rhs = (y * s.double()).sum(), (x1.double() * up.double()).sum()
    # y carries one bf16 rounding (2^-9 relative per element); the sums cancel, so bound by the sum of magnitudes
    assert abs(float(lhs - rhs)) <= 2 ** -8 * float((y.abs() * s.abs().double()).sum())


def test_tn_bf16_transposed_output():
    """Linear weight gradient written as bf16 by the GEMM epilogue (D[n][m] = sum_k A[k][m] B[k][n], m contiguous)."""
    import torch

    from disentangle_mlp_b200 import ops

    torch.manual_seed(0)
    m, n, k = 1024, 256, 64
    a = torch.randn(k, m, device="cuda").bfloat16()
    b = torch.randn(k, n, device="cuda").bfloat16()
    out = torch.zeros(n, m, device="cuda", dtype=torch.bfloat16)
    ops.gemm(ops.GEMM_TN, a, b, m, n, k, out=out, accumulate=False, ldd_m=1, ldd_n=m)
    ref = b.float().t() @ a.float()
    assert float((out.float() - ref).norm() / ref.norm()) < 4e-3

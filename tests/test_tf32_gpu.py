"""GPU: the TF32 precision mode (north-star: per-layer agreement <= 1e-3 for TF32).  fp32 operands, tcgen05.mma.kind::tf32,
fp32 accumulation and output, against torch in TRUE fp32 (allow_tf32 off) on the layer shapes of the three networks:
every dense GEMM form and every 5x5 convolution form (forward, input-gradient = transposed convolution, weight gradient),
fed the fp32 operands the reference's layers see."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    from disentangle_mlp_b200 import ops as o

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


@pytest.mark.parametrize("m,n,k", [(64, 2048, 16384), (64, 128, 2048), (64, 16384, 128), (256, 2048, 16384), (16, 2048, 1024)])
def test_linear_gemms_tf32(ops, m, n, k):
    """nn.Linear forward (NT), input-gradient (NN) and weight-gradient (TN) -- model.py:460-471, 490, 402-408"""
    torch.manual_seed(m + n)
    x = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    bias = torch.randn(n, device="cuda")
    y = ops.gemm_tf32(ops.GEMM_NT, x, w, m, n, k, bias=bias)
    assert rel(y, x @ w.t() + bias) < TOL
    dy = torch.randn(m, n, device="cuda")
    dx = ops.gemm_tf32(ops.GEMM_NN, dy, w, m, k, n)
    assert rel(dx, dy @ w) < TOL
    dw = torch.empty(n, k, device="cuda")
    ops.gemm_tf32(ops.GEMM_TN, x, dy, k, n, m, out=dw, ldd_m=1, ldd_n=k)  # D[m = in-feature][n = out-feature] -> dw[out][in]
    assert rel(dw, dy.t() @ x) < TOL
    if k >= 2048:  # split-K with fp32 reduce-add
        splits = 8
        y2 = ops.gemm_tf32(ops.GEMM_NT, x, w, m, n, k, accumulate=True, splits=splits)
        assert rel(y2, x @ w.t()) < TOL


@pytest.mark.parametrize("batch,hs,cs,cb,stride", [(16, 16, 128, 64, 2), (16, 8, 256, 128, 2), (16, 8, 256, 256, 2),
                                                   (8, 32, 128, 32, 2), (64, 16, 256, 128, 2), (3, 8, 256, 256, 2)])
def test_conv_layers_tf32(ops, batch, hs, cs, cb, stride):
    """5x5 / pad 2 Conv2d and ConvTranspose2d of the three networks (model.py:388-399, 449-457, 495-504) in TF32:
    forward, transposed direction and weight gradient, on fp32 NHWC tensors."""
    torch.manual_seed(cs + cb)
    dev = "cuda"
    w = torch.randn(cs, cb, 5, 5, device=dev) * 0.05
    bias_s, bias_b = torch.randn(cs, device=dev), torch.randn(cb, device=dev)
    g = ops.geom(batch, hs, hs, cs, cb, stride)
    wd, wu = ops.pack_conv_weights_f32(w)
    big = torch.randn(batch, hs * stride, hs * stride, cb, device=dev)
    small = torch.randn(batch, hs, hs, cs, device=dev)
    nchw = lambda t: t.permute(0, 3, 1, 2)  # noqa: E731
    y = ops.conv_down_tf32(g, big, wd, bias_s)
    assert rel(nchw(y), F.conv2d(nchw(big), w, bias_s, stride=stride, padding=2)) < TOL
    z = ops.conv_up_tf32(g, small, wu, bias_b)
    assert rel(nchw(z), F.conv_transpose2d(nchw(small), w, bias_b, stride=stride, padding=2, output_padding=stride - 1)) < TOL
    wr = w.clone().requires_grad_(True)
    F.conv2d(nchw(big), wr, None, stride=stride, padding=2).backward(nchw(small).contiguous())
    dw = ops.conv_wgrad_tf32(g, small, big)
    assert rel(dw, wr.grad) < TOL
    # and the gain over bf16: the same layer through the bf16 path is ~4x further from fp32
    wd16, _, _ = ops.pack_conv_weights(w, cs, cb)
    y16 = ops.conv_down(g, big.bfloat16(), wd16, bias_s)
    assert rel(nchw(y16), F.conv2d(nchw(big), w, bias_s, stride=stride, padding=2)) > 2 * rel(
        nchw(y), F.conv2d(nchw(big), w, bias_s, stride=stride, padding=2))

"""Tensor-level wrappers around the C ABI (include/dm_b200.h).

Every function takes CUDA torch tensors, passes raw device pointers plus the current stream to
libdm_b200.so and returns torch tensors.  torch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import GEMM_NN, GEMM_NT, GEMM_TN, BnFuse, ConvGeom, GemmDesc

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
COL_K = 80  # stored columns of the 3-channel im2col matrix (75 valid); the GEMM's K boxes zero-fill beyond
BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int | None:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "dm ops need contiguous CUDA tensors"
    return t.data_ptr()


def geom(batch, hs, ws, cs, cb, stride) -> ConvGeom:
    return ConvGeom(batch, hs, ws, cs, hs * stride, ws * stride, cb, stride)


# ------------------------------------------------------------------------------------------ GEMM-class
class BnSite:
    """Everything the kernel that PRODUCES a pre-BatchNorm tensor needs to also produce its statistics and finalize
    them (dm_bn_fuse): slot scratch, the layer's parameters / buffers, and the outputs scale_shift / mean_invstd
    ([groups,2,c]) that the consumer (bn_apply_act) and the backward pass read."""

    def __init__(self, scratch, groups, rows, c, gamma, beta, running_mean, running_var, nbt, momentum=0.1, eps=1e-5):
        dev = scratch.device
        self.scratch, self.groups, self.rows, self.c = scratch, int(groups), int(rows), int(c)
        self.gamma, self.beta, self.running_mean, self.running_var, self.nbt = gamma, beta, running_mean, running_var, nbt
        self.momentum, self.eps = momentum, eps
        self.scale_shift = torch.empty((self.groups, 2, c), dtype=F32, device=dev)
        self.mean_invstd = torch.empty((self.groups, 2, c), dtype=F32, device=dev)

    def struct(self):
        return BnFuse(_p(self.scratch), self.groups, self.rows, _p(self.gamma), _p(self.beta), _p(self.running_mean),
                      _p(self.running_var), _p(self.nbt), self.momentum, self.eps, _p(self.scale_shift),
                      _p(self.mean_invstd))


def _bn_fuse(bn):
    """bn: None or a BnSite -> dm_bn_fuse (or None)"""
    return None if bn is None else bn.struct()


def gemm(layout, a, b, m, n, k, *, out=None, out_dtype=F32, accumulate=False, bias=None, splits=1,
         lda=None, ldb=None, ldd_m=None, ldd_n=1, m_store=0, n_store=0, k_alg=0, bn=None):
    """D = op(A) op(B) with bf16 operands; see dm_gemm_desc.  bn: a BnSite = fused BatchNorm statistics + finalize."""
    assert a.dtype == BF16 and b.dtype == BF16
    if out is None:
        out = (torch.zeros if accumulate else torch.empty)((m_store or m, n_store or n), dtype=out_dtype,
                                                            device=a.device)
    if lda is None:
        lda = a.stride(0)
    if ldb is None:
        ldb = b.stride(0)
    if ldd_m is None:
        ldd_m = out.stride(0) if out.dim() >= 2 else 1
    d = GemmDesc(layout, m, n, k, _p(a), lda, _p(b), ldb, _p(out), ldd_m, ldd_n, int(out.dtype == F32),
                 int(accumulate), _p(bias), m_store, n_store, splits, k_alg, _bn_fuse(bn) or BnFuse())
    _lib.check(_lib.load().dm_gemm_bf16(C.byref(d), _stream()), "dm_gemm_bf16")
    return out


def _bn_ref(bn):
    f = _bn_fuse(bn)
    return None if f is None else C.byref(f)


def down_paired(g_cb, stride) -> bool:
    """Layers whose stride-2 convolution runs with two filter columns per k-block (dm_conv_down_paired)."""
    return stride == 2 and g_cb == 32 and os.environ.get("DM_DOWN_PAIR", "1") != "0"


def pack_down_pairs(w_down, cs, cb, out=None):
    """w_down [25][cs][cb = 32] -> w_pair [15][cs][64] (see dm_pack_down_pairs)."""
    if out is None:
        out = torch.empty((15, cs, 2 * cb), dtype=BF16, device=w_down.device)
    _lib.check(_lib.load().dm_pack_down_pairs(_p(w_down), cs, cb, _p(out), _stream()), "dm_pack_down_pairs")
    return out


def conv_down(g: ConvGeom, big, w_down, bias=None, out=None, bn=None):
    if out is None:
        out = torch.empty((g.batch, g.hs, g.ws, g.cs), dtype=BF16, device=big.device)
    if w_down.shape[0] == 15:  # paired pack
        _lib.check(_lib.load().dm_conv_down_paired(C.byref(g), _p(big), _p(w_down), _p(bias), _p(out), _bn_ref(bn),
                                                   _stream()), "dm_conv_down_paired")
        return out
    _lib.check(_lib.load().dm_conv_down(C.byref(g), _p(big), _p(w_down), _p(bias), _p(out), _bn_ref(bn), _stream()),
               "dm_conv_down")
    return out


def up_merged(g_cb, stride) -> bool:
    """Layers whose transposed convolution runs in the phase-merged form (dm_conv_up_merged)."""
    return stride == 2 and g_cb == 32 and os.environ.get("DM_UP_MERGE", "1") != "0"


def pack_up_merged(w_up, cs, cb, out=None):
    """w_up [25][cb][cs] -> w_upm [9][4*cb][cs] (see dm_pack_up_merged)."""
    if out is None:
        out = torch.empty((9, 4 * cb, cs), dtype=BF16, device=w_up.device)
    _lib.check(_lib.load().dm_pack_up_merged(_p(w_up), cs, cb, _p(out), _stream()), "dm_pack_up_merged")
    return out


def conv_up(g: ConvGeom, small, w_up, bias=None, out=None, out_f32=False, bn=None):
    if w_up.shape[0] == 9:  # phase-merged pack
        assert not out_f32
        if out is None:
            out = torch.empty((g.batch, g.hb, g.wb, g.cb), dtype=BF16, device=small.device)
        _lib.check(_lib.load().dm_conv_up_merged(C.byref(g), _p(small), _p(w_up), _p(bias), _p(out), _bn_ref(bn),
                                                 _stream()), "dm_conv_up_merged")
        return out
    if out is None:
        out = torch.empty((g.batch, g.hb, g.wb, g.cb), dtype=F32 if out_f32 else BF16, device=small.device)
    _lib.check(_lib.load().dm_conv_up(C.byref(g), _p(small), _p(w_up), _p(bias), _p(out), int(out.dtype == F32),
                                      _bn_ref(bn), _stream()), "dm_conv_up")
    return out


# 0 (default): conv weight gradients are reduced into the tap-major packed layout [25][cs][cb] by bulk tensor
# reductions (TMA) and unpacked once; 1: scattered 4-byte atomics straight into dw[cs][cb][5][5] (slower, kept for A/B)
WGRAD_DIRECT = os.environ.get("DM_WGRAD_DIRECT", "0") != "0"


def conv_wgrad_packed(g: ConvGeom, small, big, dw_packed):
    """dw_packed[25][cs][cb] (fp32, tap-major) += small^T * shifted(big)"""
    assert dw_packed.dtype == F32 and dw_packed.numel() == 25 * g.cs * g.cb
    _lib.check(_lib.load().dm_conv_wgrad(C.byref(g), _p(small), _p(big), _p(dw_packed), 0, _stream()), "dm_conv_wgrad")
    return dw_packed


def unpack_conv_grad(dw_packed, cs, cb, dw, accumulate=True):
    """dw[cs][cb][5][5] (+)= dw_packed; dw_packed is re-zeroed."""
    _lib.check(_lib.load().dm_unpack_conv_grad(_p(dw_packed), cs, cb, int(accumulate), _p(dw), _stream()),
               "dm_unpack_conv_grad")
    return dw


def conv_wgrad(g: ConvGeom, small, big, dw, packed=None, direct=None):
    """dw[cs][cb][5][5] (fp32) += small^T * shifted(big).
    direct: reduce straight into dw; otherwise through `packed` = zeroed [25][cs][cb] scratch (kept zeroed)."""
    if WGRAD_DIRECT if direct is None else direct:
        assert dw.dtype == F32 and dw.numel() == 25 * g.cs * g.cb
        _lib.check(_lib.load().dm_conv_wgrad(C.byref(g), _p(small), _p(big), _p(dw), 1, _stream()), "dm_conv_wgrad")
        return dw
    if packed is None:
        packed = torch.zeros((25, g.cs, g.cb), dtype=F32, device=dw.device)
    conv_wgrad_packed(g, small, big, packed)
    return unpack_conv_grad(packed, g.cs, g.cb, dw, True)


# ------------------------------------------------------------------------------------------ TF32 precision mode
def gemm_tf32(layout, a, b, m, n, k, *, out=None, accumulate=False, bias=None, splits=1, ldd_m=None, ldd_n=1,
              m_store=0, n_store=0):
    """D = op(A) op(B), fp32 operands read by the tensor cores as TF32, fp32 accumulate / output (dm_gemm_tf32)."""
    assert a.dtype == F32 and b.dtype == F32
    if out is None:
        out = (torch.zeros if accumulate else torch.empty)((m_store or m, n_store or n), dtype=F32, device=a.device)
    if ldd_m is None:
        ldd_m = out.stride(0) if out.dim() >= 2 else 1
    d = GemmDesc(layout, m, n, k, _p(a), a.stride(0), _p(b), b.stride(0), _p(out), ldd_m, ldd_n, 1, int(accumulate),
                 _p(bias), m_store, n_store, splits, 0, BnFuse())
    _lib.check(_lib.load().dm_gemm_tf32(C.byref(d), _stream()), "dm_gemm_tf32")
    return out


def pack_conv_weights_f32(w):
    """fp32 [cs][cb][5][5] -> fp32 (w_down [25][cs][cb], w_up [25][cb][cs]): the TF32 path's operand packs."""
    cs, cb = w.shape[0], w.shape[1]
    w_down = w.permute(2, 3, 0, 1).reshape(25, cs, cb).contiguous()
    w_up = w.permute(2, 3, 1, 0).reshape(25, cb, cs).contiguous()
    return w_down, w_up


def conv_down_tf32(g: ConvGeom, big, w_down, bias=None):
    assert big.dtype == F32 and w_down.dtype == F32
    out = torch.empty((g.batch, g.hs, g.ws, g.cs), dtype=F32, device=big.device)
    _lib.check(_lib.load().dm_conv_down_tf32(C.byref(g), _p(big), _p(w_down), _p(bias), _p(out), _stream()),
               "dm_conv_down_tf32")
    return out


def conv_up_tf32(g: ConvGeom, small, w_up, bias=None):
    assert small.dtype == F32 and w_up.dtype == F32
    out = torch.empty((g.batch, g.hb, g.wb, g.cb), dtype=F32, device=small.device)
    _lib.check(_lib.load().dm_conv_up_tf32(C.byref(g), _p(small), _p(w_up), _p(bias), _p(out), _stream()),
               "dm_conv_up_tf32")
    return out


def conv_wgrad_tf32(g: ConvGeom, small, big, dw_packed=None):
    """dw_packed [25][cs][cb] fp32 (tap-major) += small^T * shifted(big); returns it viewed as [cs][cb][5][5]."""
    assert small.dtype == F32 and big.dtype == F32
    if dw_packed is None:
        dw_packed = torch.zeros((25, g.cs, g.cb), dtype=F32, device=small.device)
    _lib.check(_lib.load().dm_conv_wgrad_tf32(C.byref(g), _p(small), _p(big), _p(dw_packed), _stream()),
               "dm_conv_wgrad_tf32")
    return dw_packed.view(5, 5, g.cs, g.cb).permute(2, 3, 0, 1)


def profile_enable(on: bool):
    _lib.load().dm_profile_enable(int(on))


def profile_read():
    """(total GEMM-kernel ms, algorithmic FLOPs, launches) since the previous read; synchronises."""
    ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(_lib.load().dm_profile_read(C.byref(ms), C.byref(fl), C.byref(n)), "dm_profile_read")
    return ms.value, fl.value, n.value


def last_plan():
    grid = (C.c_int * 3)()
    smem, stages = C.c_int(), C.c_int()
    _lib.load().dm_debug_last_plan(grid, C.byref(smem), C.byref(stages))
    return tuple(grid), smem.value, stages.value


def pack_conv_weights(w, cs, cb, want_down=True, want_up=True, want_col=False, out=None):
    """fp32 [cs][cb][5][5] -> (w_down [25][cs][cb], w_up [25][cb_pad][cs], w_col [cs][128]) bf16.
    out: a previously returned tuple to refresh in place."""
    dev = w.device
    cb_pad = max(16, (cb + 15) // 16 * 16)
    if out is not None:
        w_down, w_up, w_col = out
    else:
        w_down = torch.empty((25, cs, cb), dtype=BF16, device=dev) if want_down else None
        w_up = torch.empty((25, cb_pad, cs), dtype=BF16, device=dev) if want_up else None
        w_col = torch.empty((cs, 128), dtype=BF16, device=dev) if want_col else None
    merged = want_up and up_merged(cb, 2)
    w_up_std = w_up
    if merged:  # conv_up consumes the phase-merged pack; the standard one is only an intermediate
        w_up_std = torch.empty((25, cb_pad, cs), dtype=BF16, device=dev)
        if w_up is None or w_up.shape[0] != 9:
            w_up = torch.empty((9, 4 * cb, cs), dtype=BF16, device=dev)
    _lib.check(_lib.load().dm_pack_conv_weights(_p(w), cs, cb, _p(w_down), _p(w_up_std), _p(w_col), _stream()),
               "dm_pack_conv_weights")
    if merged:
        pack_up_merged(w_up_std, cs, cb, out=w_up)
    return w_down, w_up, w_col


# ------------------------------------------------------------------------------------------ HBM-bound
def bn_parts(rows, c) -> int:
    return int(_lib.load().dm_bn_parts(rows, c))


def bn_scratch(c, groups, device):
    """A call site's slot scratch for the BatchNorm partial sums: zero on entry AND on exit of every use (the consumer
    kernel's last block clears it), so it is allocated and zeroed once."""
    return torch.zeros(int(_lib.load().dm_bn_scratch_floats(c, groups)), dtype=F32, device=device)


def bn_slots() -> int:
    return int(_lib.load().dm_bn_slots())


def bn_stats(y, c, site: BnSite):
    """Producer + finalize for a tensor already in memory ([groups*rows, c]): fills site.scale_shift / mean_invstd,
    updates the running statistics, leaves the scratch zeroed."""
    f = site.struct()
    _lib.check(_lib.load().dm_bn_stats(_p(y), int(y.dtype == F32), c, C.byref(f), _stream()), "dm_bn_stats")
    return site.scale_shift, site.mean_invstd


def bn_apply_act(y, rows, c, scale_shift, act, slope=0.2, out=None, groups=1):
    """out = act(y * scale + shift) with given constants ([groups,2,c] or [2,c]); `rows` per group."""
    if out is None:
        out = torch.empty(y.shape, dtype=BF16, device=y.device)
    _lib.check(_lib.load().dm_bn_apply_act(_p(y), int(y.dtype == F32), rows, c, _p(scale_shift), act, slope, _p(out),
                                           groups, _stream()), "dm_bn_apply_act")
    return out


BN1D_MAX_ROWS = 256  # at most this many rows: single-launch kernels, no scratch


def bn_forward(y, rows, c, gamma, beta, running_mean, running_var, nbt, act, slope=0.2, momentum=0.1, eps=1e-5, out=None,
               groups=1, scratch=None):
    """Training-mode BatchNorm + activation (dm_bn_forward).  `y` holds `groups` stacked batches of `rows` rows each.
    Returns (out, scale_shift [groups,2,c], mean_invstd [groups,2,c]) ([2,c] each when groups == 1)."""
    if out is None:
        out = torch.empty(y.shape, dtype=BF16, device=y.device)
    if scratch is None and rows > BN1D_MAX_ROWS:
        scratch = bn_scratch(c, groups, y.device)
    scale_shift = torch.empty((groups, 2, c), dtype=F32, device=y.device)
    mean_invstd = torch.empty((groups, 2, c), dtype=F32, device=y.device)
    _lib.check(_lib.load().dm_bn_forward(_p(y), int(y.dtype == F32), rows, c, _p(gamma), _p(beta), _p(running_mean),
                                         _p(running_var), _p(nbt), momentum, eps, act, slope, _p(scratch),
                                         _p(scale_shift), _p(mean_invstd), _p(out), groups, _stream()), "dm_bn_forward")
    if groups == 1:
        return out, scale_shift[0], mean_invstd[0]
    return out, scale_shift, mean_invstd


def bn_backward(dout, y, rows, c, scale_shift, mean_invstd, act, slope=0.2, dgamma=None, dbeta=None, out=None, groups=1,
                scratch=None):
    """`rows` per group; with groups > 1 dout / y / out are the stacked tensors and scale_shift / mean_invstd are
    [groups,2,c]."""
    assert dout.dtype == BF16
    dy = torch.empty(y.shape, dtype=BF16, device=y.device) if out is None else out
    if scratch is None and rows > BN1D_MAX_ROWS:
        scratch = bn_scratch(c, groups, y.device)
    _lib.check(_lib.load().dm_bn_backward(_p(dout), _p(y), int(y.dtype == F32), rows, c, _p(scale_shift),
                                          _p(mean_invstd), act, slope, _p(scratch), _p(dy), _p(dgamma), _p(dbeta),
                                          groups, _stream()), "dm_bn_backward")
    return dy


def linear_pair_forward(x0, x1, w0, w1, b0, b1):
    """(x0 @ w0^T + b0, x1 @ w1^T + b1) in one launch: x bf16 [rows, k], w bf16 [n, k] -> fp32 [rows, n] each."""
    rows, k = x0.shape
    n = w0.shape[0]
    out = torch.empty((2, rows, n), dtype=F32, device=x0.device)
    _lib.check(_lib.load().dm_linear_pair_forward(_p(x0), _p(x1), _p(w0), _p(w1), _p(b0), _p(b1), rows, n, k, _p(out[0]),
                                                  _p(out[1]), _stream()), "dm_linear_pair_forward")
    return out[0], out[1]


def linear_pair_backward(d0, d1, x0, x1, w0, w1, dw0=None, dw1=None, db0=None, db1=None):
    """Backward of linear_pair_forward: d fp32 [rows, n] -> (dx0, dx1) bf16 [rows, k]; dw / db accumulated when given."""
    rows, k = x0.shape
    n = w0.shape[0]
    assert d0.dtype == F32 and d1.dtype == F32
    dx = torch.empty((2, rows, k), dtype=BF16, device=x0.device)
    _lib.check(_lib.load().dm_linear_pair_backward(_p(d0), _p(d1), _p(x0), _p(x1), _p(w0), _p(w1), rows, n, k, _p(dx[0]),
                                                   _p(dx[1]), _p(dw0), _p(dw1), _p(db0), _p(db1), _stream()),
               "dm_linear_pair_backward")
    return dx[0], dx[1]


def bn1d_forward_cols(acc, col0, c, pre_bias, gamma, beta, running_mean, running_var, nbt, act, slope=0.2, momentum=0.1,
                      eps=1e-5):
    """BatchNorm1d + activation of columns [col0, col0 + c) of the fp32 matrix acc [rows, ld] (dm_bn1d_forward).
    Returns (out bf16 [rows, c], scale_shift [2,c], mean_invstd [2,c]) -- constants for the STORED (bias-less) values."""
    rows, ld = acc.shape
    assert acc.dtype == F32 and acc.is_contiguous() and rows <= BN1D_MAX_ROWS
    out = torch.empty((rows, c), dtype=BF16, device=acc.device)
    ss = torch.empty((2, c), dtype=F32, device=acc.device)
    mi = torch.empty((2, c), dtype=F32, device=acc.device)
    _lib.check(_lib.load().dm_bn1d_forward(acc.data_ptr() + 4 * col0, ld, rows, c, _p(pre_bias), _p(gamma), _p(beta),
                                           _p(running_mean), _p(running_var), _p(nbt), momentum, eps, act, slope, _p(ss),
                                           _p(mi), _p(out), _stream()), "dm_bn1d_forward")
    return out, ss, mi


def bn1d_backward_cols(dout, acc, col0, c, scale_shift, mean_invstd, act, slope, dy, dgamma=None, dbeta=None):
    """Backward of bn1d_forward_cols: dout bf16 [rows, c] dense; dy bf16 [rows, ld_dy] receives columns [col0, col0+c)."""
    rows, ld = acc.shape
    assert dout.dtype == BF16 and dout.is_contiguous() and dy.dtype == BF16 and dy.is_contiguous()
    _lib.check(_lib.load().dm_bn1d_backward(_p(dout), acc.data_ptr() + 4 * col0, ld, rows, c, _p(scale_shift),
                                            _p(mean_invstd), act, slope, dy.data_ptr() + 2 * col0, dy.shape[1], _p(dgamma),
                                            _p(dbeta), _stream()), "dm_bn1d_backward")
    return dy


def bias_act(acc, rows, c, bias, act, slope=0.2, want_f32=True, want_bf16=True):
    out_f32 = torch.empty((rows, c), dtype=F32, device=acc.device) if want_f32 else None
    out_bf16 = torch.empty((rows, c), dtype=BF16, device=acc.device) if want_bf16 else None
    _lib.check(_lib.load().dm_bias_act(_p(acc), rows, c, _p(bias), act, slope, _p(out_f32), _p(out_bf16), _stream()),
               "dm_bias_act")
    return out_f32, out_bf16


def act_backward(dout, out, rows, c, act, slope, colsum):
    dpre = torch.empty((rows, c), dtype=BF16, device=dout.device)
    partials = torch.empty((bn_parts(rows, c), c), dtype=F32, device=dout.device)
    _lib.check(_lib.load().dm_act_backward(_p(dout), _p(out), rows, c, act, slope, _p(dpre), _p(partials), _p(colsum),
                                           _stream()), "dm_act_backward")
    return dpre


def colsum(x, rows, c, out):
    partials = torch.empty((bn_parts(rows, c), c), dtype=F32, device=x.device)
    _lib.check(_lib.load().dm_colsum(_p(x), int(x.dtype == F32), rows, c, _p(partials), _p(out), _stream()), "dm_colsum")
    return out


def im2col3(x_nchw, stride, out=None):
    b, ch, h, w = x_nchw.shape
    assert ch == 3 and x_nchw.dtype == F32
    if out is None:
        out = torch.empty((b * (h // stride) * (w // stride), COL_K), dtype=BF16, device=x_nchw.device)
    _lib.check(_lib.load().dm_im2col3(_p(x_nchw), b, h, w, stride, _p(out), _stream()), "dm_im2col3")
    return out


PIM_H, PIM_W, PIM_C = 68, 72, 4  # padded image: bf16 [b][68][72][4] (see dm_pad_image3)


def pim_empty(batch, device):
    """Uninitialised padded-image buffer for `batch` 64x64 images: a [batch,68,72,4] view of an allocation with the 64
    elements of slack the weight-gradient GEMM's 16-pixel windows may read (and discard) past the last image."""
    flat = torch.empty(int(_lib.load().dm_pim_elems(batch)), dtype=BF16, device=device)
    return flat[:batch * PIM_H * PIM_W * PIM_C].view(batch, PIM_H, PIM_W, PIM_C)


def pad_image3(x, pim=None, want_nchw=False):
    """image batch -> padded bf16 image (the operand of the 3-channel GEMMs).  x: fp32 NCHW [b,3,64,64] in [-1,1], or
    uint8 NHWC [b,64,64,3] (the reference's ToTensor + Normalize(.5,.5) is fused in; want_nchw then also returns the
    normalised fp32 NCHW image).  Returns pim, or (pim, x_nchw)."""
    u8 = x.dtype == torch.uint8
    b = x.shape[0]
    assert (tuple(x.shape[1:]) == (64, 64, 3)) if u8 else (tuple(x.shape[1:]) == (3, 64, 64) and x.dtype == F32)
    if pim is None:
        pim = pim_empty(b, x.device)
    nchw = torch.empty((b, 3, 64, 64), dtype=F32, device=x.device) if (u8 and want_nchw) else None
    _lib.check(_lib.load().dm_pad_image3(_p(x), int(u8), b, _p(pim), _p(nchw), _stream()), "dm_pad_image3")
    return (pim, nchw if u8 else x) if want_nchw else pim


def conv3_cols(cs, stride):
    """GEMM columns of a 3-image-channel layer: cs, or 2*cs = (pixel of the output pair, channel) for stride 1"""
    return 2 * cs if stride == 1 else cs


def pack_conv3_weights(w, stride, out=None):
    """fp32 [cs][3][5][5] -> bf16 window pack [5][cs or 2*cs][32] (dm_pack_conv3_weights)"""
    cs = w.shape[0]
    if out is None:
        out = torch.empty((5, conv3_cols(cs, stride), 32), dtype=BF16, device=w.device)
    _lib.check(_lib.load().dm_pack_conv3_weights(_p(w), cs, stride, _p(out), _stream()), "dm_pack_conv3_weights")
    return out


def conv3_fwd(g: ConvGeom, pim, w_win, bias=None, out=None, bn=None):
    """bf16 NHWC [b,hs,ws,cs] = conv5x5(3-channel image, W) + bias from the padded image (stride g.stride)."""
    if out is None:
        out = torch.empty((g.batch, g.hs, g.ws, g.cs), dtype=BF16, device=pim.device)
    _lib.check(_lib.load().dm_conv3_fwd(C.byref(g), _p(pim), _p(w_win), _p(bias), _p(out), _bn_ref(bn), _stream()),
               "dm_conv3_fwd")
    return out


def conv3_wgrad(g: ConvGeom, pim, small, dw, scratch=None):
    """dw[cs][3][5][5] (fp32) += sum_pixels small[b,h,w,cs] x 5x5 image patch.  scratch: zeroed fp32
    [5][conv3_cols(cs, stride)][64] window gradient buffer (kept zeroed by the unpack kernel)."""
    if scratch is None:
        scratch = torch.zeros((5, conv3_cols(g.cs, g.stride), 64), dtype=F32, device=pim.device)
    assert scratch.numel() == 5 * conv3_cols(g.cs, g.stride) * 64
    _lib.check(_lib.load().dm_conv3_wgrad(C.byref(g), _p(pim), _p(small), _p(scratch), _stream()), "dm_conv3_wgrad")
    _lib.check(_lib.load().dm_unpack_conv3_grad(_p(scratch), g.cs, g.stride, _p(dw), _stream()), "dm_unpack_conv3_grad")
    return dw


def nhwc3_to_nchw(src, batch, h, w, apply_tanh, pim=None):
    dst = torch.empty((batch, 3, h, w), dtype=F32, device=src.device)
    _lib.check(_lib.load().dm_nhwc3_to_nchw(_p(src), batch, h * w, int(apply_tanh), _p(dst), _p(pim), _stream()),
               "dm_nhwc3_to_nchw")
    return dst


def tanh_backward(dout, out, bias_grad=None, pim=None, want_dy=True):
    """dy = dout * (1 - out^2): as fp32 NCHW (want_dy) and / or as a padded bf16 image (pim)."""
    b, ch, h, w = out.shape
    dy = torch.empty_like(out) if want_dy else None
    _lib.check(_lib.load().dm_tanh_backward(_p(dout), _p(out), b, h * w, _p(dy), _p(bias_grad), _p(pim), _stream()),
               "dm_tanh_backward")
    return dy


def transpose(src, batch, rows, cols, out=None):
    dst = torch.empty((batch, cols, rows), dtype=BF16, device=src.device) if out is None else out
    _lib.check(_lib.load().dm_transpose_bf16(_p(src), batch, rows, cols, _p(dst), _stream()), "dm_transpose_bf16")
    return dst


def cast_bf16(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    _lib.check(_lib.load().dm_cast_bf16(_p(src), src.numel(), _p(dst), _stream()), "dm_cast_bf16")
    return dst


def reparam_forward(mu, logvar, eps, out_bf16=None):
    z = torch.empty_like(mu)
    zb = torch.empty(mu.shape, dtype=BF16, device=mu.device) if out_bf16 is None else out_bf16
    _lib.check(_lib.load().dm_reparam_forward(_p(mu), _p(logvar), _p(eps), mu.numel(), _p(z), _p(zb), _stream()),
               "dm_reparam_forward")
    return z, zb


def reparam_backward(dz, logvar, eps, dmu_ext=None, dlogvar_ext=None):
    dmu = torch.empty(logvar.shape, dtype=BF16, device=logvar.device)
    dlv = torch.empty(logvar.shape, dtype=BF16, device=logvar.device)
    dmu32 = torch.empty_like(logvar)
    dlv32 = torch.empty_like(logvar)
    _lib.check(_lib.load().dm_reparam_backward(_p(dz), _p(logvar), _p(eps), _p(dmu_ext), _p(dlogvar_ext),
                                               logvar.numel(), _p(dmu), _p(dlv), _p(dmu32), _p(dlv32), _stream()),
               "dm_reparam_backward")
    return dmu, dlv, dmu32, dlv32


def head_forward(feat, w, b):
    rows, k = feat.shape
    prob = torch.empty((rows,), dtype=F32, device=feat.device)
    _lib.check(_lib.load().dm_head_forward(_p(feat), rows, k, _p(w), _p(b), _p(prob), _stream()), "dm_head_forward")
    return prob


def head_backward(dprob, prob, feat, dfeat_ext, w, dw, db):
    rows, k = feat.shape
    dfeat = torch.empty_like(feat)
    _lib.check(_lib.load().dm_head_backward(_p(dprob), _p(prob), _p(feat), _p(dfeat_ext), rows, k, _p(w), _p(dfeat),
                                            _p(dw), _p(db), _stream()), "dm_head_backward")
    return dfeat


def mse_sum(a, b, loss, wloss=1.0, grad=None, wgrad=1.0, accumulate=False):
    _lib.check(_lib.load().dm_mse_sum(_p(a), _p(b), a.numel(), wloss, _p(loss), wgrad, int(accumulate), _p(grad),
                                      _stream()), "dm_mse_sum")


def kl(mu, logvar, loss, w=1.0, dmu=None, dlogvar=None, accumulate=False):
    _lib.check(_lib.load().dm_kl(_p(mu), _p(logvar), mu.numel(), w, _p(loss), int(accumulate), _p(dmu), _p(dlogvar),
                                 _stream()), "dm_kl")


def bce_const(p, target, loss, w=1.0, n_total=None, dprob=None, accumulate=False, stat=None):
    """`target`: python float, or a 1-element CUDA tensor (read on the device: CUDA-graph replay)."""
    n = p.numel()
    tdev = target if isinstance(target, torch.Tensor) else None
    _lib.check(_lib.load().dm_bce_const(_p(p), n, float(n_total or n), 0.0 if tdev is not None else float(target),
                                        _p(tdev), w, _p(loss), int(accumulate), _p(dprob), _p(stat), _stream()),
               "dm_bce_const")


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0, shadow=None, step_dev=None, count_step=True,
              enable=None, step_offset=None):
    """step_dev: int32 CUDA tensor holding the step count; incremented on the device (unless count_step is False: a
    later segment of the same optimizer step) and used instead of `step`.  g: fp32 or bf16.
    enable: int32 CUDA scalar gating a DEFERRED update (no-op when 0); step_offset: an EARLY update that runs before
    the step counter of its optimizer step has been incremented (uses step_dev + step_offset).  Either implies
    count_step=False."""
    assert g.dtype in (F32, BF16) and g.numel() == p.numel()
    if enable is not None or step_offset is not None:
        _lib.check(_lib.load().dm_adam_step_gated(_p(p), _p(g), int(g.dtype == BF16), _p(m), _p(v), p.numel(), lr, beta1,
                                                  beta2, eps, _p(step_dev), int(step_offset or 0), grad_scale, _p(shadow),
                                                  _p(enable), _stream()), "dm_adam_step_gated")
        return
    _lib.check(_lib.load().dm_adam_step_ex(_p(p), _p(g), int(g.dtype == BF16), _p(m), _p(v), p.numel(), lr, beta1, beta2,
                                           eps, step, _p(step_dev), int(count_step), grad_scale, _p(shadow), _stream()),
               "dm_adam_step")

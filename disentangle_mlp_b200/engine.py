"""Forward / backward schedules of the three networks as explicit sequences of libdm_b200 kernel calls.

No autograd in here: each `*_forward` returns its outputs plus a `Saved` record, each `*_backward` consumes
that record and ACCUMULATES parameter gradients into the fp32 tensors it is handed (zero them first for a
fresh gradient).  model.py wraps these in torch.autograd.Function for the reference training loops;
steps.py calls them directly for the fused steps.

Layouts: activations bf16 NHWC; the 16384-wide Linear layers are defined on the NCHW flatten order
(model.py:516-517, 540-543, 412-413), so one small bf16 transpose sits on either side of them.
Parameters are addressed by their reference state_dict names.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from types import SimpleNamespace

import torch

from . import ops
from .ops import ACT_LEAKY, ACT_RELU, BF16, F32, GEMM_NN, GEMM_NT, GEMM_TN

BN_EPS, BN_MOMENTUM, LEAKY = 1e-5, 0.1, 0.2


# ------------------------------------------------------------------------------------------ operand cache
class OperandCache:
    """bf16 GEMM operands derived from the fp32 master parameters; rebuilt when a parameter's version
    counter or storage changes (optimizer step, load_state_dict, weights_init), or after invalidate()."""

    def __init__(self):
        self._c = {}
        self.wgrad_scratch = {}  # (weight name, device) -> zeroed [25][cs][cb] fp32 packed-gradient scratch
        self.bn_scratch = {}  # (site, c, groups, device) -> BatchNorm slot scratch, zero between uses (ops.bn_scratch)
        self.static = False  # True: never rebuild implicitly (CUDA-graph mode); refresh() does it in place

    def get(self, key, param, builder):
        ent = self._c.get(key)
        ver = (param.data_ptr(), param._version)
        if ent is not None and (self.static or ent[0] == ver):
            return ent[1]
        val = builder(param.detach())
        self._c[key] = (ver, val)
        return val

    def invalidate(self):
        self._c.clear()


def _lin_w(cache, name, p):
    """bf16 copy of a Linear weight: a view of the optimizer's bf16 shadow when there is one, else a cast."""
    views = getattr(cache, "lin_views", None)
    if views is not None:
        return views[name + ".weight"]
    return cache.get(("lin", name), p, lambda w: ops.cast_bf16(w.contiguous()))


FUSE_HEADS = os.environ.get("DM_FUSE_HEADS", "1") != "0"
PAIR_HEADS = os.environ.get("DM_PAIR_HEADS", "0") != "0"  # measured slower than the two tensor-core launches (4.46 vs 4.42 ms): off
HEADS = ("x_to_mu", "x_to_logvar")


def _stacked_rows(a, b):
    """[a; b] as ONE matrix when b's rows follow a's in the same buffer (views of a flat parameter / gradient buffer),
    else None."""
    if a is None or b is None or a.dtype != b.dtype or a.shape != b.shape or not (a.is_contiguous() and b.is_contiguous()):
        return None
    if b.data_ptr() != a.data_ptr() + a.numel() * a.element_size():
        return None
    return a.as_strided((2 * a.shape[0], a.shape[1]), (a.shape[1], 1))


def _heads_weight(cache, P):
    """The two heads' 16384 -> 2048 weights as one [4096, 16384] bf16 operand (fused trainers: adjacent in the optimizer's
    bf16 shadow), or None (module path: separate casts)."""
    views = getattr(cache, "lin_views", None)
    if not FUSE_HEADS or views is None:
        return None
    return _stacked_rows(views.get(HEADS[0] + ".0.weight"), views.get(HEADS[1] + ".0.weight"))


CONV3_STRIDE = {"convs.0": 1, "features.0": 2, "deconv4": 1}  # the three layers on the 3-channel image side


def pack3(w, stride, out=None):
    """operand packs of a 3-image-channel layer [cs][3][5][5]: (None, w_up kw-folded [25][16][cs] for the 32 -> 3
    transposed direction, w_win for the GEMMs over the padded image (ops.pack_conv3_weights))"""
    cs = w.shape[0]
    w = w.contiguous()
    if out is None:
        out = (None, torch.empty((25, 16, cs), dtype=BF16, device=w.device), None)
    ops.pack_conv_weights(w, cs, 3, False, True, False, out=(None, out[1], None))
    return (None, out[1], ops.pack_conv3_weights(w, stride, out=out[2]))


def _conv_pack(cache, name, p, cs, cb):
    """(w_down [25][cs][cb], w_up [25][cb_pad][cs], w_pair [15][cs][64] for cb = 32 else None); for the three 3-channel
    layers (None, w_up, w_win): pack3"""
    static = getattr(cache, "static_packs", None)
    if static is not None:  # fused trainers: persistent buffers refreshed in place after every optimizer step
        return static[name]
    if cb == 3:
        return cache.get(("conv", name), p, lambda w: pack3(w, CONV3_STRIDE[name]))

    def build(w):
        w_down, w_up, _ = ops.pack_conv_weights(w.contiguous(), cs, cb, True, True, False)
        return (w_down, w_up, ops.pack_down_pairs(w_down, cs, cb) if ops.down_paired(cb, 2) else None)

    return cache.get(("conv", name), p, build)


def _down_w(packs):
    """the conv_down operand of a _conv_pack() triple: the paired pack [15][cs][64] where the layer has one"""
    return packs[0] if packs[2] is None else packs[2]


# ------------------------------------------------------------------------------------------ BatchNorm helper
@dataclass
class BNState:
    y: torch.Tensor          # pre-normalisation tensor (bf16 or fp32), viewed as [rows, c]
    rows: int
    c: int
    scale_shift: torch.Tensor
    mean_invstd: torch.Tensor
    act: int


FUSE_BN_STATS = os.environ.get("DM_BN_FUSE_GEMM", "1") != "0"  # A/B: statistics in the producing GEMM's epilogue


def _site_scratch(cache, site, c, groups, dev):
    """The persistent slot scratch of one BatchNorm call site (zero between uses: the consumer kernel clears it)."""
    key = (site, c, groups, str(dev))
    sc = cache.bn_scratch.get(key)
    if sc is None:
        sc = cache.bn_scratch[key] = ops.bn_scratch(c, groups, dev)
    return sc


def bn_fuse(cache, P, B, prefix, c, groups, rows_per_group, dev, training=True, tile_rows=None):
    """ops.BnSite for the GEMM that writes the pre-BatchNorm tensor of layer `prefix`, or None when the statistics
    cannot ride in its epilogue (eval mode, passes that do not cover whole 128-row tiles, A/B switch).
    rows_per_group: rows of the tensor per stacked pass; tile_rows: GEMM rows per pass if different (a transposed
    convolution tiles its SMALL side: a quarter of the output pixels).
    Pass the result to BOTH the GEMM (bn=...) and bn_act_forward (fused=...)."""
    if not (training and FUSE_BN_STATS) or (tile_rows or rows_per_group) % 128 != 0:
        return None
    return ops.BnSite(_site_scratch(cache, prefix + "/f", c, groups, dev), groups, rows_per_group, c,
                      P[prefix + ".weight"].detach(), P[prefix + ".bias"].detach(), B[prefix + ".running_mean"],
                      B[prefix + ".running_var"], B[prefix + ".num_batches_tracked"], BN_MOMENTUM, BN_EPS)


def bn_act_forward(y, rows, c, P, B, prefix, act, training=True, groups=1, cache=None, fused=None):
    """BatchNorm (batch statistics, running-stat update) + activation. P: params, B: buffers.

    groups > 1: `y` holds `groups` independent batches stacked along rows (several forward passes of the same
    network pushed through each GEMM together); statistics, normalisation and the running-stat updates are done
    per group, in order -- exactly what separate forward calls would do -- by ONE set of kernel launches
    (blockIdx.z = group).  fused: the bn_fuse(...) site the producing GEMM was given (its epilogue produced the
    statistics and its last CTA finalized them).  Returns (out, [BNState per group])."""
    gamma, beta = P[prefix + ".weight"], P[prefix + ".bias"]
    rm, rv, nbt = B[prefix + ".running_mean"], B[prefix + ".running_var"], B[prefix + ".num_batches_tracked"]
    rg = rows // groups
    y2 = y.view(rows, c)
    out = torch.empty(y.shape, dtype=BF16, device=y.device)
    o2 = out.view(rows, c)
    if training:
        if fused is not None:  # the GEMM's last CTA has already finalized: constants are in fused.scale_shift
            ss, mi = fused.scale_shift, fused.mean_invstd
            ops.bn_apply_act(y2, rg, c, ss, act, LEAKY, out=o2, groups=groups)
        else:
            sc = None
            if rg > ops.BN1D_MAX_ROWS and cache is not None:
                sc = _site_scratch(cache, prefix + "/f", c, groups, y.device)
            _, ss, mi = ops.bn_forward(y2, rg, c, gamma.detach(), beta.detach(), rm, rv, nbt, act, LEAKY, BN_MOMENTUM,
                                       BN_EPS, out=o2, groups=groups, scratch=sc)
            if groups == 1:
                ss, mi = ss.unsqueeze(0), mi.unsqueeze(0)
        return out, [BNState(y2[g * rg:(g + 1) * rg], rg, c, ss[g], mi[g], act) for g in range(groups)]
    states = []
    for g in range(groups):  # inference statistics (the reference scripts never call .eval(); kept for completeness)
        ys = y2[g * rg:(g + 1) * rg]
        invstd = torch.rsqrt(rv + BN_EPS)
        sc = gamma.detach() * invstd
        ss = torch.stack([sc, beta.detach() - rm * sc]).contiguous()
        mi = torch.stack([rm, invstd]).contiguous()
        ops.bn_apply_act(ys, rg, c, ss, act, LEAKY, out=o2[g * rg:(g + 1) * rg])
        states.append(BNState(ys, rg, c, ss, mi, act))
    return out, states


def _stacked(states):
    """True if the per-group states are adjacent slices of one stacked forward (same rows, consecutive memory)."""
    s0 = states[0]
    ye = s0.y.element_size()
    for i, st in enumerate(states):
        if st.rows != s0.rows or st.c != s0.c or st.act != s0.act:
            return False
        if st.y.data_ptr() != s0.y.data_ptr() + i * s0.rows * s0.c * ye:
            return False
        if st.scale_shift.data_ptr() != s0.scale_shift.data_ptr() + i * 2 * s0.c * 4:
            return False
        if st.mean_invstd.data_ptr() != s0.mean_invstd.data_ptr() + i * 2 * s0.c * 4:
            return False
    return True


def bn_act_backward(dout, states, G, prefix, cache=None):
    """Backward of bn_act_forward over the given per-group states (dout rows stacked in the same order).
    Returns dy (bf16); accumulates dgamma / dbeta into G when present."""
    dg = G.get(prefix + ".weight") if G is not None else None
    db = G.get(prefix + ".bias") if G is not None else None
    c = states[0].c
    rows = sum(st.rows for st in states)
    d2 = dout.view(rows, c)
    dy = torch.empty(dout.shape, dtype=BF16, device=dout.device)
    y2 = dy.view(rows, c)
    def scratch(groups, rows_g):
        if rows_g <= ops.BN1D_MAX_ROWS or cache is None:
            return None
        return _site_scratch(cache, prefix + "/b", c, groups, dout.device)

    if len(states) > 1 and _stacked(states):
        st = states[0]
        ops.bn_backward(d2, st.y, st.rows, st.c, st.scale_shift, st.mean_invstd, st.act, LEAKY, dg, db, out=y2,
                        groups=len(states), scratch=scratch(len(states), st.rows))
        return dy
    r0 = 0
    for st in states:
        ops.bn_backward(d2[r0:r0 + st.rows], st.y, st.rows, st.c, st.scale_shift, st.mean_invstd, st.act, LEAKY, dg, db,
                        out=y2[r0:r0 + st.rows], scratch=scratch(1, st.rows))
        r0 += st.rows
    return dy


def _splits_for(m_tiles, n_tiles, kblocks, target=None):
    """Split-K factor of a skinny Linear GEMM: enough CTAs to keep the weight stream at HBM speed."""
    if target is None:
        target = int(os.environ.get("DM_LIN_CTAS", "296"))
    s = max(1, min(kblocks, target // max(1, m_tiles * n_tiles)))
    return s


def linear_forward(x_bf16, w_bf16, bias, batch, n_out, k_in, out_dtype=F32):
    """y[batch, n_out] = x @ w^T + bias, fp32 accumulate; split-K (atomic fp32) when the grid would be tiny."""
    n_tiles = (n_out + 127) // 128
    m_tiles = (batch + 127) // 128
    splits = _splits_for(m_tiles, n_tiles, k_in // 64 if k_in >= 64 else 1)
    if out_dtype != F32:
        splits = 1
    return ops.gemm(GEMM_NT, x_bf16, w_bf16, batch, n_out, k_in, out_dtype=out_dtype, accumulate=splits > 1,
                    bias=bias, splits=splits)


def linear_dgrad(dy_bf16, w_bf16, batch, n_out, k_in, out_dtype=BF16, out=None):
    """dx[batch, k_in] = dy[batch, n_out] @ w[n_out, k_in]"""
    n_tiles = (k_in + 127) // 128
    m_tiles = (batch + 127) // 128
    splits = _splits_for(m_tiles, n_tiles, max(1, n_out // 64)) if (out_dtype == F32) else 1
    acc = splits > 1 or out is not None
    return ops.gemm(GEMM_NN, dy_bf16, w_bf16, batch, k_in, n_out, out=out, out_dtype=out_dtype, accumulate=acc,
                    splits=splits)


def linear_wgrad(dy_bf16, x_bf16, batch, n_out, k_in, dw, overwrite=False):
    """dw[n_out, k_in] (+)= dy^T @ x.  Computed as D[m = in-feature, n = out-feature] so that a warp's 32 rows are
    contiguous floats of dw (coalesced); `overwrite` stores instead of accumulating (dw need not be zeroed)."""
    assert dw.dtype == F32 or overwrite, "a bf16 weight-gradient buffer is written once per phase (overwrite)"
    ops.gemm(GEMM_TN, x_bf16, dy_bf16, k_in, n_out, batch, out=dw, accumulate=not overwrite, ldd_m=1, ldd_n=k_in)


class WgradSide:
    """Convolution weight gradients on a side stream.  Inside one network's backward pass the weight gradient of a layer
    and the rest of the chain (input gradient -> BatchNorm backward -> next layer) only share their inputs, and nothing
    reads the weight gradient before the optimizer: with `stream` set (the fused trainers do, DM_WGRAD_STREAM=1) every
    conv weight-gradient launch forks off the current stream, and `join()` -- called at the end of each backward
    pass -- brings it back.  The tail of each GEMM (last tiles, epilogue, drain: ~5 us of a 20-40 us launch at batch
    64) then overlaps another kernel's work instead of idling the SMs.  Works the same inside a CUDA-graph capture
    (fork / join become graph edges)."""
    stream = None
    big_stream = None  # second side stream: the 16384x2048 Linear weight gradients (HBM-latency-bound, like the
                       # input-gradient GEMM they now run beside) and, behind them, the early Adam update of those weights
    keep = []      # operands of the in-flight side-stream launches: alive until join()
    pending = set()

    @classmethod
    def run(cls, fn, *args, big=False):
        s = cls.big_stream if big else cls.stream
        if s is None:
            fn(*args)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        s.wait_event(ev)
        with torch.cuda.stream(s):
            fn(*args)
        cls.keep.append(args)
        cls.pending.add(s)

    @classmethod
    def join(cls):
        for s in cls.pending:
            ev = torch.cuda.Event()
            ev.record(s)
            torch.cuda.current_stream().wait_event(ev)
        cls.pending.clear()
        cls.keep.clear()


def conv_wgrad(g, small, big, dw, cache, name):
    """Conv / ConvT weight gradient through the tap-major packed scratch (persistent per weight, kept zeroed)."""
    WgradSide.run(_conv_wgrad, g, small, big, dw, cache, name)


def _conv_wgrad(g, small, big, dw, cache, name):
    packed = getattr(cache, "packed_grads", None)
    if packed is not None and (name + ".weight") in packed:
        # fused trainers: the flat gradient buffer holds this weight's gradient in the packed layout already
        gp = packed[name + ".weight"]
        assert gp.data_ptr() == dw.data_ptr(), "packed gradient view does not alias the gradient handed in"
        ops.conv_wgrad_packed(g, small, big, gp)
        return
    if ops.WGRAD_DIRECT:
        ops.conv_wgrad(g, small, big, dw)
        return
    ws = cache.wgrad_scratch
    key = (name, dw.device)
    if key not in ws:
        ws[key] = torch.zeros((25, g.cs, g.cb), dtype=F32, device=dw.device)
    ops.conv_wgrad(g, small, big, dw, ws[key])


def conv3_wgrad(g, pim, small, dw, cache, name):
    """Weight gradient of a 3-image-channel layer through its persistent window-layout scratch (kept zeroed)."""
    WgradSide.run(_conv3_wgrad, g, pim, small, dw, cache, name)


def _conv3_wgrad(g, pim, small, dw, cache, name):
    ws = cache.wgrad_scratch
    key = (name, "win", str(dw.device))
    if key not in ws:
        ws[key] = torch.zeros((5, ops.conv3_cols(g.cs, g.stride), 64), dtype=F32, device=dw.device)
    ops.conv3_wgrad(g, pim, small, dw, ws[key])


# ------------------------------------------------------------------------------------------ Discriminator
def discriminator_forward(x, P, B, cache: OperandCache, training=True, groups=1, pim=None, before_linear=None):
    """Discriminator_celeba.forward (model.py:410-416). x: fp32 NCHW [b,3,64,64], or None with pim = the batch as a
    padded bf16 image (ops.pad_image3 / the decoder's pim output).  Returns prob [b], feat [b,2048].
    groups > 1: the batch stacks `groups` separate batches (e.g. real | fake); every GEMM processes them together while
    BatchNorm treats each group as its own forward pass (see bn_act_forward)."""
    S = SimpleNamespace(groups=groups)
    S.pim = ops.pad_image3(x) if pim is None else pim
    b = S.b = S.pim.shape[0]
    _, _, ww1 = _conv_pack(cache, "convs.0", P["convs.0.weight"], 32, 3)
    dev, bg = S.pim.device, b // groups
    f1 = bn_fuse(cache, P, B, "convs.1", 32, groups, bg * 4096, dev, training)
    raw1 = ops.conv3_fwd(ops.geom(b, 64, 64, 32, 3, 1), S.pim, ww1, P["convs.0.bias"].detach(), bn=f1)
    S.a1, S.bn1 = bn_act_forward(raw1, b * 4096, 32, P, B, "convs.1", ACT_LEAKY, training, groups, cache, f1)
    g2 = ops.geom(b, 32, 32, 128, 32, 2)
    wd2 = _down_w(_conv_pack(cache, "convs.3", P["convs.3.weight"], 128, 32))
    f2 = bn_fuse(cache, P, B, "convs.4", 128, groups, bg * 1024, dev, training)
    raw2 = ops.conv_down(g2, S.a1, wd2, P["convs.3.bias"].detach(), bn=f2)
    S.a2, S.bn2 = bn_act_forward(raw2, b * 1024, 128, P, B, "convs.4", ACT_LEAKY, training, groups, cache, f2)
    g3 = ops.geom(b, 16, 16, 256, 128, 2)
    wd3, _, _ = _conv_pack(cache, "convs.6", P["convs.6.weight"], 256, 128)
    f3 = bn_fuse(cache, P, B, "convs.7", 256, groups, bg * 256, dev, training)
    raw3 = ops.conv_down(g3, S.a2, wd3, P["convs.6.bias"].detach(), bn=f3)
    S.a3, S.bn3 = bn_act_forward(raw3, b * 256, 256, P, B, "convs.7", ACT_LEAKY, training, groups, cache, f3)
    g4 = ops.geom(b, 8, 8, 256, 256, 2)
    wd4, _, _ = _conv_pack(cache, "convs.9", P["convs.9.weight"], 256, 256)
    f4 = bn_fuse(cache, P, B, "convs.10", 256, groups, bg * 64, dev, training)
    raw4 = ops.conv_down(g4, S.a3, wd4, P["convs.9.bias"].detach(), bn=f4)
    a4, S.bn4 = bn_act_forward(raw4, b * 64, 256, P, B, "convs.10", ACT_LEAKY, training, groups, cache, f4)
    S.flat = ops.transpose(a4, b, 64, 256)  # NHWC [b,64,256] -> NCHW flatten order [b,256*64]
    if before_linear:  # data parallel, sharded Adam: the all-gather of the updated bf16 weight must have landed
        before_linear()
    wl = _lin_w(cache, "lth_features.0", P["lth_features.0.weight"])
    acc = linear_forward(S.flat, wl, None, b, 2048, 16384)
    S.feat, _ = ops.bias_act(acc, b, 2048, P["lth_features.0.bias"].detach(), ACT_LEAKY, LEAKY, True, False)
    S.prob = ops.head_forward(S.feat, P["sigmoid_output.0.weight"].detach().contiguous(),
                              P["sigmoid_output.0.bias"].detach())
    return S.prob, S.feat, S


def _slice_disc(S, g0, g1):
    """View of a discriminator forward record restricted to groups [g0, g1) (batch-major tensors: contiguous)."""
    if g0 == 0 and g1 == S.groups:
        return S
    bg = S.b // S.groups
    i0, i1 = g0 * bg, g1 * bg
    V = SimpleNamespace(b=i1 - i0, groups=g1 - g0)
    V.pim = S.pim[i0:i1]
    V.a1, V.a2, V.a3 = S.a1[i0:i1], S.a2[i0:i1], S.a3[i0:i1]
    V.flat, V.feat, V.prob = S.flat[i0:i1], S.feat[i0:i1], S.prob[i0:i1]
    V.bn1, V.bn2, V.bn3, V.bn4 = S.bn1[g0:g1], S.bn2[g0:g1], S.bn3[g0:g1], S.bn4[g0:g1]
    return V


def discriminator_backward(S, dprob, dfeat, P, G, cache: OperandCache, need_dx=True, need_wgrad=True,
                           overwrite_big=False, grad_ready=None, group_range=None, linear_done=None):
    """Backward of discriminator_forward. dprob [b] / dfeat [b,2048] fp32 (either may be None).
    G: dict name -> fp32 grad tensor (accumulated) or None. Returns dx fp32 NCHW or None.
    group_range=(g0, g1): back-propagate only through those groups of a grouped forward (b = their images).
    linear_done: called once the 16384x2048 Linear is back-propagated -- its weight gradient is final and the weight is
    not read again in this pass (its optimizer update may start, see encoder_backward's heads_done)."""
    if group_range is not None:
        S = _slice_disc(S, *group_range)
    b = S.b
    dev = S.feat.device
    if dprob is None:
        dprob = torch.zeros(b, dtype=F32, device=dev)
    wg = G if need_wgrad else None
    wo = P["sigmoid_output.0.weight"].detach().contiguous()
    dfeat_t = ops.head_backward(dprob.contiguous(), S.prob, S.feat, None if dfeat is None else dfeat.contiguous(), wo,
                                wg["sigmoid_output.0.weight"] if wg else None,
                                wg["sigmoid_output.0.bias"] if wg else None)
    colsum = wg["lth_features.0.bias"] if wg else None
    dpre = ops.act_backward(dfeat_t, S.feat, b, 2048, ACT_LEAKY, LEAKY, colsum)
    wl = _lin_w(cache, "lth_features.0", P["lth_features.0.weight"])
    if wg:
        def big_wgrad(dpre_, flat_, dw_):
            linear_wgrad(dpre_, flat_, b, 2048, 16384, dw_, overwrite_big)
            if grad_ready:  # the 33.5 M-element gradient is final: its all-reduce can overlap the conv backward below
                grad_ready("lth_features.0.weight")
        WgradSide.run(big_wgrad, dpre, S.flat, wg["lth_features.0.weight"], big=True)
    dflat = linear_dgrad(dpre, wl, b, 2048, 16384)
    if linear_done:
        linear_done()
    da4 = ops.transpose(dflat, b, 256, 64)  # back to NHWC [b,64,256]
    # conv 4
    dr4 = bn_act_backward(da4, S.bn4, wg, "convs.10", cache)
    g4 = ops.geom(b, 8, 8, 256, 256, 2)
    _, wu4, _ = _conv_pack(cache, "convs.9", P["convs.9.weight"], 256, 256)
    if wg:
        conv_wgrad(g4, dr4, S.a3, wg["convs.9.weight"], cache, "convs.9")
    da3 = ops.conv_up(g4, dr4, wu4)
    # conv 3
    dr3 = bn_act_backward(da3, S.bn3, wg, "convs.7", cache)
    g3 = ops.geom(b, 16, 16, 256, 128, 2)
    _, wu3, _ = _conv_pack(cache, "convs.6", P["convs.6.weight"], 256, 128)
    if wg:
        conv_wgrad(g3, dr3, S.a2, wg["convs.6.weight"], cache, "convs.6")
    da2 = ops.conv_up(g3, dr3, wu3)
    # conv 2
    dr2 = bn_act_backward(da2, S.bn2, wg, "convs.4", cache)
    g2 = ops.geom(b, 32, 32, 128, 32, 2)
    _, wu2, _ = _conv_pack(cache, "convs.3", P["convs.3.weight"], 128, 32)
    if wg:
        conv_wgrad(g2, dr2, S.a1, wg["convs.3.weight"], cache, "convs.3")
    da1 = ops.conv_up(g2, dr2, wu2)
    # conv 1 (3 input channels: GEMMs over the padded image)
    dr1 = bn_act_backward(da1, S.bn1, wg, "convs.1", cache)
    g1 = ops.geom(b, 64, 64, 32, 3, 1)
    if wg:
        conv3_wgrad(g1, S.pim, dr1, wg["convs.0.weight"], cache, "convs.0")
    if not need_dx:
        WgradSide.join()
        return None
    _, wu1, _ = _conv_pack(cache, "convs.0", P["convs.0.weight"], 32, 3)
    dx_nhwc = ops.conv_up(g1, dr1, wu1, out_f32=True)
    dx = ops.nhwc3_to_nchw(dx_nhwc, b, 64, 64, False)
    WgradSide.join()
    return dx


# ------------------------------------------------------------------------------------------ Encoder
def encoder_forward(x, P, B, cache: OperandCache, training=True, pim=None, before_heads=None):
    """VAE.encode (model.py:511-522). x: fp32 NCHW, or None with pim = the batch as a padded bf16 image.
    Returns mu, logvar fp32 [b,128]."""
    S = SimpleNamespace()
    S.pim = ops.pad_image3(x) if pim is None else pim
    b = S.b = S.pim.shape[0]
    _, _, ww1 = _conv_pack(cache, "features.0", P["features.0.weight"], 64, 3)
    dev = S.pim.device
    f1 = bn_fuse(cache, P, B, "features.1", 64, 1, b * 1024, dev, training)
    raw1 = ops.conv3_fwd(ops.geom(b, 32, 32, 64, 3, 2), S.pim, ww1, P["features.0.bias"].detach(), bn=f1)
    S.a1, S.bn1 = bn_act_forward(raw1, b * 1024, 64, P, B, "features.1", ACT_RELU, training, 1, cache, f1)
    g2 = ops.geom(b, 16, 16, 128, 64, 2)
    wd2, _, _ = _conv_pack(cache, "features.3", P["features.3.weight"], 128, 64)
    f2 = bn_fuse(cache, P, B, "features.4", 128, 1, b * 256, dev, training)
    raw2 = ops.conv_down(g2, S.a1, wd2, P["features.3.bias"].detach(), bn=f2)
    S.a2, S.bn2 = bn_act_forward(raw2, b * 256, 128, P, B, "features.4", ACT_RELU, training, 1, cache, f2)
    g3 = ops.geom(b, 8, 8, 256, 128, 2)
    wd3, _, _ = _conv_pack(cache, "features.6", P["features.6.weight"], 256, 128)
    f3 = bn_fuse(cache, P, B, "features.7", 256, 1, b * 64, dev, training)
    raw3 = ops.conv_down(g3, S.a2, wd3, P["features.6.bias"].detach(), bn=f3)
    a3, S.bn3 = bn_act_forward(raw3, b * 64, 256, P, B, "features.7", ACT_RELU, training, 1, cache, f3)
    S.flat = ops.transpose(a3, b, 64, 256)
    if before_heads:  # (see discriminator_forward)
        before_heads()
    outs = []
    S.heads = {}
    wcat = _heads_weight(cache, P) if (training and b <= ops.BN1D_MAX_ROWS) else None
    S.acc_cat = None
    if wcat is not None:
        # both heads' first Linear as ONE GEMM over their stacked weights (N = 4096): one pass over the activations, half
        # the launches; the per-head BatchNorm1d reads its column block and adds the head's Linear bias itself
        S.acc_cat = linear_forward(S.flat, wcat, None, b, 4096, 16384)
    for i, head in enumerate(HEADS):
        if wcat is not None:
            h1, ss, mi = ops.bn1d_forward_cols(S.acc_cat, i * 2048, 2048, P[head + ".0.bias"].detach(),
                                               P[head + ".1.weight"].detach(), P[head + ".1.bias"].detach(),
                                               B[head + ".1.running_mean"], B[head + ".1.running_var"],
                                               B[head + ".1.num_batches_tracked"], ACT_RELU, LEAKY, BN_MOMENTUM, BN_EPS)
            bn = SimpleNamespace(scale_shift=ss, mean_invstd=mi)
        else:
            w0 = _lin_w(cache, head + ".0", P[head + ".0.weight"])
            acc = linear_forward(S.flat, w0, P[head + ".0.bias"].detach(), b, 2048, 16384)
            h1, bn = bn_act_forward(acc, b, 2048, P, B, head + ".1", ACT_RELU, training, 1, cache)
        S.heads[head] = SimpleNamespace(h1=h1, bn=bn)
    w3 = [_lin_w(cache, head + ".3", P[head + ".3.weight"]) for head in HEADS]
    b3 = [P[head + ".3.bias"].detach() for head in HEADS]
    if PAIR_HEADS:  # both heads' Linear(2048, 128) in one SIMT launch (17 MFLOP each: a tensor-core launch is all overhead)
        outs = ops.linear_pair_forward(S.heads[HEADS[0]].h1, S.heads[HEADS[1]].h1, w3[0], w3[1], b3[0], b3[1])
    else:
        outs = [linear_forward(S.heads[head].h1, w3[i], b3[i], b, 128, 2048) for i, head in enumerate(HEADS)]
    return outs[0], outs[1], S


def encoder_backward(S, dmu, dlogvar, P, G, cache: OperandCache, need_wgrad=True, overwrite_big=False,
                     grad_ready=None, heads_done=None):
    """dmu / dlogvar: fp32 [b,128] gradients w.r.t. the encoder outputs.
    heads_done: called once both heads are back-propagated -- the two 16384x2048 weight gradients are final and the
    weights themselves are not read again in this pass (their optimizer update may start)."""
    b = S.b
    dev = S.flat.device
    wg = G if need_wgrad else None
    fused = getattr(S, "acc_cat", None) is not None
    wcat = _heads_weight(cache, P) if fused else None
    gcat = _stacked_rows(wg[HEADS[0] + ".0.weight"], wg[HEADS[1] + ".0.weight"]) if (fused and wg) else None
    assert not fused or (wcat is not None and (gcat is not None or not wg)), \
        "fused heads: the two 16384x2048 weights / gradients must be adjacent in the flat buffers"
    dflat = None if fused else torch.zeros((b, 16384), dtype=F32, device=dev)
    dacc_cat = torch.empty((b, 4096), dtype=BF16, device=dev) if fused else None
    ds = [torch.zeros((b, 128), dtype=F32, device=dev) if d is None else d.contiguous() for d in (dmu, dlogvar)]
    pair = PAIR_HEADS and b <= 256
    if pair:  # both heads' Linear(2048, 128): input gradient, weight gradient and bias gradient in two SIMT launches
        w3 = [_lin_w(cache, head + ".3", P[head + ".3.weight"]) for head in HEADS]
        dh1s = ops.linear_pair_backward(ds[0], ds[1], S.heads[HEADS[0]].h1, S.heads[HEADS[1]].h1, w3[0], w3[1],
                                        *([wg[h + ".3.weight"] for h in HEADS] + [wg[h + ".3.bias"] for h in HEADS]
                                          if wg else [None] * 4))
    for i, (head, d) in enumerate(zip(HEADS, ds)):
        H = S.heads[head]
        if pair:
            dh1 = dh1s[i]
        else:
            d16 = ops.cast_bf16(d)
            if wg:
                ops.colsum(d, b, 128, wg[head + ".3.bias"])
                linear_wgrad(d16, H.h1, b, 128, 2048, wg[head + ".3.weight"])
            w3 = _lin_w(cache, head + ".3", P[head + ".3.weight"])
            dh1 = linear_dgrad(d16, w3, b, 128, 2048)
        if fused:
            ops.bn1d_backward_cols(dh1, S.acc_cat, i * 2048, 2048, H.bn.scale_shift, H.bn.mean_invstd, ACT_RELU, LEAKY,
                                   dacc_cat, wg[head + ".1.weight"] if wg else None, wg[head + ".1.bias"] if wg else None)
            continue
        dacc = bn_act_backward(dh1, H.bn, wg, head + ".1", cache)
        w0 = _lin_w(cache, head + ".0", P[head + ".0.weight"])
        if wg:
            def big_wgrad(dacc_, flat_, dw_, head_=head):
                linear_wgrad(dacc_, flat_, b, 2048, 16384, dw_, overwrite_big)
                if grad_ready:
                    grad_ready(head_ + ".0.weight")
            WgradSide.run(big_wgrad, dacc, S.flat, wg[head + ".0.weight"], big=True)
        linear_dgrad(dacc, w0, b, 2048, 16384, out_dtype=F32, out=dflat)
    if fused:
        # one weight-gradient GEMM into the two (adjacent) gradients, one input-gradient GEMM over K = 4096 written
        # straight in bf16 (no fp32 accumulation buffer, no cast)
        if wg:
            def big_wgrad(dacc_, flat_, dw_):
                linear_wgrad(dacc_, flat_, b, 4096, 16384, dw_, overwrite_big)
                if grad_ready:
                    for head in HEADS:
                        grad_ready(head + ".0.weight")
            WgradSide.run(big_wgrad, dacc_cat, S.flat, gcat, big=True)
        dflat16 = linear_dgrad(dacc_cat, wcat, b, 4096, 16384)
    else:
        dflat16 = None
    if heads_done:
        heads_done()
    da3 = ops.transpose(dflat16 if fused else ops.cast_bf16(dflat), b, 256, 64)
    dr3 = bn_act_backward(da3, S.bn3, wg, "features.7", cache)
    g3 = ops.geom(b, 8, 8, 256, 128, 2)
    _, wu3, _ = _conv_pack(cache, "features.6", P["features.6.weight"], 256, 128)
    if wg:
        conv_wgrad(g3, dr3, S.a2, wg["features.6.weight"], cache, "features.6")
    da2 = ops.conv_up(g3, dr3, wu3)
    dr2 = bn_act_backward(da2, S.bn2, wg, "features.4", cache)
    g2 = ops.geom(b, 16, 16, 128, 64, 2)
    _, wu2, _ = _conv_pack(cache, "features.3", P["features.3.weight"], 128, 64)
    if wg:
        conv_wgrad(g2, dr2, S.a1, wg["features.3.weight"], cache, "features.3")
    da1 = ops.conv_up(g2, dr2, wu2)
    dr1 = bn_act_backward(da1, S.bn1, wg, "features.1", cache)
    if wg:
        conv3_wgrad(ops.geom(b, 32, 32, 64, 3, 2), S.pim, dr1, wg["features.0.weight"], cache, "features.0")
    WgradSide.join()
    return None  # the encoder input is data: no input gradient on this path


# ------------------------------------------------------------------------------------------ Decoder
def decoder_forward(code, P, B, cache: OperandCache, training=True, pim_out=None, groups=1):
    """VAE.decode / Generator_celeba.forward (model.py:537-566, 363-378). code: fp32 or bf16 [b,128].
    Returns recon fp32 NCHW [b,3,64,64]; pim_out (optional [b,68,72,8] bf16 buffer) also receives it as a padded image,
    the form in which the discriminator reads it.
    groups > 1: `code` stacks that many separate decode() calls (e.g. decode(noise) | decode(z)); every GEMM processes
    them together, BatchNorm treats each group as its own forward pass, in order (see bn_act_forward)."""
    b = code.shape[0]
    bg = b // groups
    S = SimpleNamespace(b=b, groups=groups)
    S.code16 = code if code.dtype == BF16 else ops.cast_bf16(code.contiguous())
    wp = _lin_w(cache, "preprocess.0", P["preprocess.0.weight"])
    acc = linear_forward(S.code16, wp, P["preprocess.0.bias"].detach(), b, 16384, 128)
    h, S.bn0 = bn_act_forward(acc, b, 16384, P, B, "preprocess.1", ACT_RELU, training, groups, cache)
    S.h0 = ops.transpose(h, b, 256, 64)  # NCHW flatten order -> NHWC [b,8,8,256]
    dev = code.device
    g1 = ops.geom(b, 8, 8, 256, 256, 2)
    _, wu1, _ = _conv_pack(cache, "deconv1", P["deconv1.weight"], 256, 256)
    f1 = bn_fuse(cache, P, B, "act1.0", 256, groups, bg * 256, dev, training, tile_rows=bg * 64)
    raw1 = ops.conv_up(g1, S.h0, wu1, P["deconv1.bias"].detach(), bn=f1)
    S.a1, S.bn1 = bn_act_forward(raw1, b * 256, 256, P, B, "act1.0", ACT_RELU, training, groups, cache, f1)
    g2 = ops.geom(b, 16, 16, 256, 128, 2)
    _, wu2, _ = _conv_pack(cache, "deconv2", P["deconv2.weight"], 256, 128)
    f2 = bn_fuse(cache, P, B, "act2.0", 128, groups, bg * 1024, dev, training, tile_rows=bg * 256)
    raw2 = ops.conv_up(g2, S.a1, wu2, P["deconv2.bias"].detach(), bn=f2)
    S.a2, S.bn2 = bn_act_forward(raw2, b * 1024, 128, P, B, "act2.0", ACT_RELU, training, groups, cache, f2)
    g3 = ops.geom(b, 32, 32, 128, 32, 2)
    _, wu3, _ = _conv_pack(cache, "deconv3", P["deconv3.weight"], 128, 32)
    f3 = bn_fuse(cache, P, B, "act3.0", 32, groups, bg * 4096, dev, training, tile_rows=bg * 1024)
    raw3 = ops.conv_up(g3, S.a2, wu3, P["deconv3.bias"].detach(), bn=f3)
    S.a3, S.bn3 = bn_act_forward(raw3, b * 4096, 32, P, B, "act3.0", ACT_RELU, training, groups, cache, f3)
    g4 = ops.geom(b, 64, 64, 32, 3, 1)
    _, wu4, _ = _conv_pack(cache, "deconv4", P["deconv4.weight"], 32, 3)
    y4 = ops.conv_up(g4, S.a3, wu4, P["deconv4.bias"].detach(), out_f32=True)  # fp32 NHWC(3)
    S.recon = ops.nhwc3_to_nchw(y4, b, 64, 64, True, pim=pim_out)
    return S.recon, S


def decoder_backward(S, drecon, P, G, cache: OperandCache, need_dcode=True, need_wgrad=True, overwrite_big=False):
    """drecon: fp32 NCHW gradient w.r.t. the decoder output (all groups of a grouped forward, stacked).
    Returns dcode fp32 [b,128] or None."""
    b = S.b
    wg = G if need_wgrad else None
    # gradient w.r.t. the pre-tanh image, written straight as a padded bf16 image: the operand of both GEMMs below
    pim4 = ops.pim_empty(b, drecon.device)
    ops.tanh_backward(drecon.contiguous(), S.recon, wg["deconv4.bias"] if wg else None, pim=pim4, want_dy=False)
    g4 = ops.geom(b, 64, 64, 32, 3, 1)
    if wg:
        conv3_wgrad(g4, pim4, S.a3, wg["deconv4.weight"], cache, "deconv4")
    _, _, ww4 = _conv_pack(cache, "deconv4", P["deconv4.weight"], 32, 3)
    da3 = ops.conv3_fwd(g4, pim4, ww4, None)  # ConvT input-gradient = conv of dy with the same weights
    dr3 = bn_act_backward(da3, S.bn3, wg, "act3.0", cache)
    g3 = ops.geom(b, 32, 32, 128, 32, 2)
    wd3 = _down_w(_conv_pack(cache, "deconv3", P["deconv3.weight"], 128, 32))
    if wg:
        conv_wgrad(g3, S.a2, dr3, wg["deconv3.weight"], cache, "deconv3")
    da2 = ops.conv_down(g3, dr3, wd3)
    dr2 = bn_act_backward(da2, S.bn2, wg, "act2.0", cache)
    g2 = ops.geom(b, 16, 16, 256, 128, 2)
    wd2, _, _ = _conv_pack(cache, "deconv2", P["deconv2.weight"], 256, 128)
    if wg:
        conv_wgrad(g2, S.a1, dr2, wg["deconv2.weight"], cache, "deconv2")
    da1 = ops.conv_down(g2, dr2, wd2)
    dr1 = bn_act_backward(da1, S.bn1, wg, "act1.0", cache)
    g1 = ops.geom(b, 8, 8, 256, 256, 2)
    wd1, _, _ = _conv_pack(cache, "deconv1", P["deconv1.weight"], 256, 256)
    if wg:
        conv_wgrad(g1, S.h0, dr1, wg["deconv1.weight"], cache, "deconv1")
    dh0 = ops.conv_down(g1, dr1, wd1)  # NHWC [b,8,8,256]
    dh = ops.transpose(dh0, b, 64, 256)  # -> [b, 256*64] flatten order
    dacc = bn_act_backward(dh, S.bn0, wg, "preprocess.1", cache)
    if wg:
        linear_wgrad(dacc, S.code16, b, 16384, 128, wg["preprocess.0.weight"], overwrite_big)
    if not need_dcode:
        WgradSide.join()
        return None
    wp = _lin_w(cache, "preprocess.0", P["preprocess.0.weight"])
    dcode = linear_dgrad(dacc, wp, b, 16384, 128, out_dtype=F32)
    WgradSide.join()
    return dcode

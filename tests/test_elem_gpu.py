"""GPU: the HBM-bound kernels through the C ABI against torch fp32 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    from disentangle_mlp_b200 import ops as o

    return o


def _bn_case(ops, rows, c, dtype, act, seed, mean_shift=0.3, use_scratch=None, reps=1, rm0=None):
    torch.manual_seed(seed)
    y = (torch.randn(rows, c, device="cuda") * 1.7 + mean_shift).to(dtype)
    gamma = torch.randn(c, device="cuda") * 0.1 + 1
    beta = torch.randn(c, device="cuda") * 0.1
    for rep in range(reps):
        rm = torch.zeros(c, device="cuda") if rm0 is None else rm0.clone()
        rv = torch.ones(c, device="cuda")
        nbt = torch.zeros((), dtype=torch.long, device="cuda")
        rm_ref, rv_ref = rm.clone(), rv.clone()
        out, ss, mi = ops.bn_forward(y, rows, c, gamma, beta, rm, rv, nbt, act, 0.2, scratch=use_scratch)
        yr = y.float().clone().requires_grad_(True)  # clone: y.float() aliases an fp32 y
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        z = F.batch_norm(yr, rm_ref, rv_ref, gr, br, training=True, momentum=0.1, eps=1e-5)
        ref = {0: z, 1: F.relu(z), 2: F.leaky_relu(z, 0.2)}[act]
        assert rel(out, ref) < 4e-3  # bf16 output rounding
        assert rel(rm, rm_ref) < 1e-4 and rel(rv, rv_ref) < 1e-4 and int(nbt) == 1
        dout = torch.randn(rows, c, device="cuda").bfloat16()
        ref.backward(dout.float())
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        dy = ops.bn_backward(dout, y, rows, c, ss, mi, act, 0.2, dg, db, scratch=use_scratch)
        assert rel(dy, yr.grad) < 6e-3
        assert rel(dg, gr.grad) < 2e-3 and rel(db, br.grad) < 2e-3
        if use_scratch is not None:  # the producer's last block leaves slots + ticket zeroed for the next use
            n0 = ops.bn_slots() * 2 * c + 4  # (behind them: the backward pass's per-group sums, overwritten per use)
            assert float(use_scratch[:n0].abs().max()) == 0.0


@pytest.mark.parametrize("rows,c,dtype,act", [(8 * 4096, 32, torch.bfloat16, 2), (4096, 256, torch.bfloat16, 1),
                                             (16, 2048, torch.float32, 1), (16, 16384, torch.float32, 1),
                                             (64, 16384, torch.float32, 1), (256, 2048, torch.float32, 1),
                                             (1000, 64, torch.bfloat16, 0), (64 * 4096, 32, torch.bfloat16, 2),
                                             (300, 2048, torch.float32, 1), (130, 256, torch.bfloat16, 1)])
def test_batchnorm_forward_backward(ops, rows, c, dtype, act):
    """dm_bn_forward / dm_bn_backward against torch fp32: rows <= 256 take the single-launch two-pass kernels, larger
    tensors the producer (slot partial sums) + consumer (finalize prologue, last block cleans up) pair."""
    _bn_case(ops, rows, c, dtype, act, seed=0)


def test_batchnorm_scratch_is_reusable(ops):
    """One persistent scratch per call site, zero on entry and on exit: three uses in a row without a memset."""
    rows, c = 4096, 128
    sc = ops.bn_scratch(c, 1, "cuda")
    _bn_case(ops, rows, c, torch.bfloat16, 2, seed=3, use_scratch=sc, reps=3)


def test_batchnorm_large_mean_small_std(ops):
    """|mean| >> std: E[y^2] - E[y]^2 in fp32 would cancel; the shifted sums (k = running_mean, here close to the
    batch mean as after a few training steps) and the two-pass small-row kernel must not."""
    torch.manual_seed(5)
    for rows, c in ((4096, 64), (64, 2048)):
        y = (torch.randn(rows, c, device="cuda") * 1e-2 + 50.0)
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        rm, rv = torch.full((c,), 49.9, device="cuda"), torch.ones(c, device="cuda")
        rm_ref, rv_ref = rm.clone(), rv.clone()
        out, ss, mi = ops.bn_forward(y, rows, c, gamma, beta, rm, rv, None, 0, 0.2)
        ref = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
        assert rel(out, ref) < 8e-3, (rows, c, rel(out, ref))
        assert rel(rv, rv_ref) < 1e-3


@pytest.mark.parametrize("kind", ["gemm", "conv_down", "conv_up", "conv_up_merged"])
def test_batchnorm_statistics_fused_into_gemm_epilogue(ops, kind):
    """The statistics ride in the epilogue of the GEMM that writes the pre-BatchNorm tensor and its last CTA finalizes
    them (dm_bn_fuse): dm_bn_apply_act with the constants it wrote must equal torch's batch_norm of the tensor the GEMM
    wrote -- for 3 stacked passes (per-pass statistics, running stats in pass order)."""
    torch.manual_seed(7)
    G, dev = 3, "cuda"
    if kind == "gemm":
        m, n, k = G * 1024, 32, 80
        a = torch.randn(m, k, device=dev).bfloat16()
        a[1024:2048] *= 2.0
        w = (torch.randn(n, 128, device=dev) * 0.1).bfloat16()
        c, rows = n, m // G
        run = lambda bn: ops.gemm(ops.GEMM_NT, a, w, m, n, k, out_dtype=torch.bfloat16,  # noqa: E731
                                  bias=torch.linspace(-1, 1, n, device=dev), bn=bn)
    else:
        b = 2 * G
        cs, cb, hs = {"conv_down": (128, 64, 16), "conv_up": (256, 128, 8), "conv_up_merged": (128, 32, 16)}[kind]
        g = ops.geom(b, hs, hs, cs, cb, 2)
        wt = torch.randn(cs, cb, 5, 5, device=dev) * 0.05
        wd, wu, _ = ops.pack_conv_weights(wt, cs, cb)
        if kind == "conv_down":
            x = torch.randn(b, 2 * hs, 2 * hs, cb, device=dev).bfloat16()
            x[2:4] *= 3.0
            c, rows = cs, (b // G) * hs * hs
            run = lambda bn: ops.conv_down(g, x, wd, torch.linspace(-1, 1, cs, device=dev), bn=bn)  # noqa: E731
        else:
            x = torch.randn(b, hs, hs, cs, device=dev).bfloat16()
            x[2:4] *= 3.0
            c, rows = cb, (b // G) * 4 * hs * hs
            assert (wu.shape[0] == 9) == (kind == "conv_up_merged")
            run = lambda bn: ops.conv_up(g, x, wu, torch.linspace(-1, 1, cb, device=dev), bn=bn)  # noqa: E731
    gamma = torch.randn(c, device=dev) * 0.1 + 1
    beta = torch.randn(c, device=dev) * 0.1
    rm, rv = torch.randn(c, device=dev) * 0.1, torch.ones(c, device=dev)
    nbt = torch.zeros((), dtype=torch.long, device=dev)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    sc = ops.bn_scratch(c, G, dev)
    for rep in range(2):  # twice: the scratch comes back zeroed
        site = ops.BnSite(sc, G, rows, c, gamma, beta, rm, rv, nbt)
        y = run(site)  # the GEMM's epilogue sums, its last CTA finalizes
        y2 = y.view(G * rows, c)
        out = ops.bn_apply_act(y2, rows, c, site.scale_shift, 2, 0.2, groups=G)
        ref = torch.cat([F.leaky_relu(F.batch_norm(y2[i * rows:(i + 1) * rows].float(), rm_ref, rv_ref, gamma, beta,
                                                   training=True, momentum=0.1, eps=1e-5), 0.2) for i in range(G)])
        assert rel(out, ref) < 5e-3, (kind, rep, rel(out, ref))
        assert rel(rm, rm_ref) < 2e-3 and rel(rv, rv_ref) < 2e-3 and int(nbt) == G * (rep + 1), (kind, rep)
        assert float(sc[:G * ops.bn_slots() * 2 * c + 4].abs().max()) == 0.0
    plain = run(None)
    assert torch.equal(plain, y)  # the fused statistics do not change what the GEMM writes


def test_batchnorm_groups(ops):
    """groups = 3 stacked passes through one set of launches == three separate BatchNorm calls in order (per-pass
    statistics, running stats updated pass by pass); backward over the sub-range of passes 1..2."""
    torch.manual_seed(2)
    rows, c, G = 2048, 64, 3
    y = (torch.randn(G * rows, c, device="cuda") * torch.tensor([1.0, 2.0, 0.5], device="cuda").repeat_interleave(rows)[:, None]
         + 0.4).bfloat16()
    gamma = torch.randn(c, device="cuda") * 0.1 + 1
    beta = torch.randn(c, device="cuda") * 0.1
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    out, ss, mi = ops.bn_forward(y, rows, c, gamma, beta, rm, rv, nbt, 2, 0.2, groups=G)
    rm_ref, rv_ref = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    yr = y.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    refs = [F.leaky_relu(F.batch_norm(yr[g * rows:(g + 1) * rows], rm_ref, rv_ref, gr, br, training=True, momentum=0.1,
                                      eps=1e-5), 0.2) for g in range(G)]
    ref = torch.cat(refs)
    assert rel(out, ref) < 4e-3
    assert rel(rm, rm_ref) < 1e-4 and rel(rv, rv_ref) < 1e-4 and int(nbt) == G
    dout = torch.randn(2 * rows, c, device="cuda").bfloat16()
    torch.cat(refs[1:]).backward(dout.float())
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dy = ops.bn_backward(dout, y[rows:], rows, c, ss[1:], mi[1:], 2, 0.2, dg, db, groups=2)
    assert rel(dy, yr.grad[rows:]) < 6e-3
    assert rel(dg, gr.grad) < 2e-3 and rel(db, br.grad) < 2e-3


@pytest.mark.parametrize("stride", [1, 2])
def test_im2col3(ops, stride):
    x = torch.rand(3, 3, 64, 64, device="cuda") * 2 - 1
    col = ops.im2col3(x, stride)
    ref = F.unfold(x, 5, padding=2, stride=stride).transpose(1, 2).reshape(-1, 75)
    assert torch.equal(col[:, :75].float(), ref.bfloat16().float())
    assert col.shape[1] == 80 and float(col[:, 75:].abs().max()) == 0.0


def _unpad(pim):
    """padded bf16 image [b,68,72,4] -> (interior as fp32 NCHW [b,3,64,64], everything else)"""
    inner = pim[:, 2:66, 2:66, :3].float().permute(0, 3, 1, 2)
    rest = pim.clone()
    rest[:, 2:66, 2:66, :3] = 0
    return inner, rest


def test_padded_image_producers(ops):
    """dm_pad_image3 (fp32 NCHW and uint8 NHWC with the loader's ToTensor + Normalize(.5,.5) fused in,
    dataloader/dataset.py:37-43), and the padded-image outputs of the decoder's tanh / tanh-backward kernels."""
    torch.manual_seed(0)
    x = torch.rand(5, 3, 64, 64, device="cuda") * 2 - 1
    pim = ops.pad_image3(x, ops.pim_empty(5, "cuda").fill_(7.0))  # (pre-filled: borders must be ZEROED by the kernel)
    inner, rest = _unpad(pim)
    assert torch.equal(inner, x.bfloat16().float()) and float(rest.abs().max()) == 0.0
    u8 = torch.randint(0, 256, (5, 64, 64, 3), dtype=torch.uint8, device="cuda")
    pim, xn = ops.pad_image3(u8, ops.pim_empty(5, "cuda").fill_(7.0), want_nchw=True)
    ref = ((u8.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5  # transforms.ToTensor() + Normalize((.5,)*3, (.5,)*3)
    assert float((xn - ref).abs().max()) < 1e-6
    inner, rest = _unpad(pim)
    assert torch.equal(inner, xn.bfloat16().float()) and float(rest.abs().max()) == 0.0
    y = torch.randn(4, 64, 64, 3, device="cuda")
    p2 = ops.pim_empty(4, "cuda").fill_(7.0)
    out = ops.nhwc3_to_nchw(y, 4, 64, 64, True, pim=p2)
    inner, rest = _unpad(p2)
    assert torch.equal(inner, out.bfloat16().float()) and float(rest.abs().max()) == 0.0
    dout = torch.randn_like(out)
    p3 = ops.pim_empty(4, "cuda").fill_(7.0)
    bg = torch.zeros(3, device="cuda")
    assert ops.tanh_backward(dout, out, bg, pim=p3, want_dy=False) is None
    dy = dout * (1 - out * out)
    inner, rest = _unpad(p3)  # (the kernel's 1 - o*o is one fma: last-bit differences before the bf16 rounding)
    assert rel(inner, dy) < 4e-3 and float((inner - dy.bfloat16().float()).abs().max()) < 0.05
    assert float(rest.abs().max()) == 0.0
    assert rel(bg, dy.sum(dim=(0, 2, 3))) < 1e-5


@pytest.mark.parametrize("batch,cs,stride", [(6, 32, 1), (6, 64, 2), (64, 32, 1), (3, 64, 2), (1, 32, 1)])
def test_conv3_window_gemms(ops, batch, cs, stride):
    """The 3-channel layers as implicit GEMMs over the padded image (overlapping TMA windows) against torch fp32 on the
    same bf16-rounded operands: forward + bias, weight gradient; with and without fused BatchNorm statistics."""
    from disentangle_mlp_b200 import engine

    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(batch + cs)
    x = torch.rand(batch, 3, 64, 64, device="cuda") * 2 - 1
    w = torch.randn(cs, 3, 5, 5, device="cuda") * 0.1
    bias = torch.randn(cs, device="cuda") * 0.1
    hs = 64 // stride
    g = ops.geom(batch, hs, hs, cs, 3, stride)
    pim = ops.pad_image3(x)
    _, _, ww = engine.pack3(w, stride)
    y = ops.conv3_fwd(g, pim, ww, bias)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias, stride=stride, padding=2)
    assert rel(y.float().permute(0, 3, 1, 2), ref) < 4e-3
    dy = torch.randn(batch, hs, hs, cs, device="cuda").bfloat16()
    dw = torch.zeros_like(w)
    sc = torch.zeros(5, ops.conv3_cols(cs, stride), 64, device="cuda")
    ops.conv3_wgrad(g, pim, dy, dw, sc)
    ops.conv3_wgrad(g, pim, dy, dw, sc)  # accumulates; the scratch comes back zeroed
    wr = w.clone().requires_grad_(True)
    F.conv2d(x.bfloat16().float(), wr, None, stride=stride, padding=2).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(dw, 2 * wr.grad) < 4e-3 and float(sc.abs().max()) == 0.0
    rows = batch * hs * hs
    if rows % 128 == 0:
        site = ops.BnSite(ops.bn_scratch(cs, 1, "cuda"), 1, rows, cs, torch.ones(cs, device="cuda"), torch.zeros(cs, device="cuda"),
                          torch.zeros(cs, device="cuda"), torch.ones(cs, device="cuda"), None)
        y2 = ops.conv3_fwd(g, pim, ww, bias, bn=site)
        assert torch.equal(y2, y)
        yf = y.float().view(rows, cs)
        assert rel(site.mean_invstd[0, 0], yf.mean(0)) < 2e-3
        assert rel(site.mean_invstd[0, 1], torch.rsqrt(yf.var(0, unbiased=False) + 1e-5)) < 2e-3


def test_layout_kernels(ops):
    x = torch.randn(5, 64, 256, device="cuda").bfloat16()
    assert torch.equal(ops.transpose(x, 5, 64, 256), x.transpose(1, 2).contiguous())
    y = torch.randn(4, 64, 64, 3, device="cuda")
    assert torch.equal(ops.nhwc3_to_nchw(y, 4, 64, 64, False), y.permute(0, 3, 1, 2).contiguous())
    assert rel(ops.nhwc3_to_nchw(y, 4, 64, 64, True), torch.tanh(y).permute(0, 3, 1, 2)) < 1e-6
    out = torch.tanh(y).permute(0, 3, 1, 2).contiguous()
    dout = torch.randn_like(out)
    bg = torch.zeros(3, device="cuda")
    dy = ops.tanh_backward(dout, out, bg)  # (fp32 NCHW output; the padded-image form is tested above)
    ref = dout * (1 - out * out)
    assert rel(dy, ref) < 1e-6 and rel(bg, ref.sum((0, 2, 3))) < 1e-4


def test_pack_conv_weights(ops, monkeypatch):
    monkeypatch.setenv("DM_UP_MERGE", "0")  # the plain [25][cb][cs] transposed-conv pack (merged form: next test)
    w = torch.randn(64, 3, 5, 5, device="cuda")
    wd, wu, wc = ops.pack_conv_weights(w, 64, 3, True, True, True)
    wb = w.bfloat16()
    assert torch.equal(wd, wb.reshape(64, 3, 25).permute(2, 0, 1).contiguous())
    # cb == 3: kw-folded layout wu[kh][kw*3 + cb][cs] = W[cs][cb][kh][kw]; everything else zero
    assert torch.equal(wu[:5, :15, :], wb.permute(2, 3, 1, 0).reshape(5, 15, 64).contiguous())
    assert float(wu[:5, 15:, :].float().abs().max()) == 0.0 and float(wu[5:].float().abs().max()) == 0.0
    w2 = torch.randn(128, 32, 5, 5, device="cuda")
    wd2, wu2, _ = ops.pack_conv_weights(w2, 128, 32)
    assert torch.equal(wu2, w2.bfloat16().reshape(128, 32, 25).permute(2, 1, 0).contiguous())
    assert torch.equal(wc[:, :75], wb.reshape(64, 75)) and float(wc[:, 75:].float().abs().max()) == 0.0


def test_pack_up_merged(ops):
    """cb == 32: conv_up consumes the phase-merged pack [9][4*cb][cs] (column = (ph*2+pw)*cb + c, tap = (dh, dw))."""
    cs, cb = 128, 32
    w = torch.randn(cs, cb, 5, 5, device="cuda")
    _, wum, _ = ops.pack_conv_weights(w, cs, cb)
    assert tuple(wum.shape) == (9, 4 * cb, cs)
    ref = torch.zeros(9, 4, cb, cs, device="cuda", dtype=torch.bfloat16)
    wb = w.bfloat16()
    for t in range(9):
        dh, dw = 1 - t // 3, 1 - t % 3
        for ph in range(2):
            for pw in range(2):
                kh, kw = ph + 2 - 2 * dh, pw + 2 - 2 * dw
                if 0 <= kh < 5 and 0 <= kw < 5:
                    ref[t, ph * 2 + pw] = wb[:, :, kh, kw].t()
    assert torch.equal(wum, ref.view(9, 4 * cb, cs))


def test_bias_act_and_backward(ops):
    acc = torch.randn(32, 2048, device="cuda")
    bias = torch.randn(2048, device="cuda")
    o32, o16 = ops.bias_act(acc, 32, 2048, bias, 2, 0.2)
    ref = F.leaky_relu(acc + bias, 0.2)
    assert rel(o32, ref) < 1e-6 and rel(o16, ref) < 4e-3
    dout = torch.randn_like(acc)
    cs = torch.zeros(2048, device="cuda")
    dpre = ops.act_backward(dout, o32, 32, 2048, 2, 0.2, cs)
    refd = dout * torch.where(ref > 0, 1.0, 0.2)
    assert rel(dpre, refd) < 4e-3 and rel(cs, refd.sum(0)) < 1e-4


def test_head_and_reparam(ops):
    feat = torch.randn(16, 2048, device="cuda")
    w = (torch.randn(1, 2048, device="cuda") * 0.02).requires_grad_(True)
    b = torch.zeros(1, device="cuda").requires_grad_(True)
    fr = feat.clone().requires_grad_(True)
    pref = torch.sigmoid(F.linear(fr, w, b)).squeeze()
    prob = ops.head_forward(feat, w.detach(), b.detach())
    assert rel(prob, pref) < 1e-5
    dprob = torch.randn(16, device="cuda")
    dfe = torch.randn(16, 2048, device="cuda")
    (pref * dprob).sum().backward()
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    dfeat = ops.head_backward(dprob, prob, feat, dfe, w.detach(), dw, db)
    assert rel(dfeat, fr.grad + dfe) < 1e-4 and rel(dw, w.grad) < 1e-4 and rel(db, b.grad) < 1e-4
    mu, lv, eps = (torch.randn(16, 128, device="cuda") for _ in range(3))
    z, z16 = ops.reparam_forward(mu, lv, eps)
    assert rel(z, mu + eps * torch.exp(0.5 * lv)) < 1e-6 and rel(z16, z) < 4e-3
    dz = torch.randn_like(mu)
    _, _, dmu, dlv = ops.reparam_backward(dz, lv, eps)
    assert rel(dmu, dz) < 1e-6 and rel(dlv, dz * eps * 0.5 * torch.exp(0.5 * lv)) < 1e-6


def test_losses(ops):
    a, b = torch.randn(8, 3, 64, 64, device="cuda"), torch.randn(8, 3, 64, 64, device="cuda")
    loss = torch.zeros((), device="cuda")
    g = torch.empty_like(a)
    ops.mse_sum(a, b, loss, 0.5, g, 0.5)
    assert abs(float(loss) - 0.5 * float(F.mse_loss(a, b, reduction="sum"))) < 1e-3 * float(loss)
    assert rel(g, a - b) < 1e-6
    mu, lv = torch.randn(8, 128, device="cuda"), torch.randn(8, 128, device="cuda")
    loss.zero_()
    dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
    ops.kl(mu, lv, loss, 25.0, dmu, dlv)
    mr, lr = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    ref = 25.0 * (-0.5 * torch.sum(1 + lr - mr.pow(2) - lr.exp()))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    assert rel(dmu, mr.grad) < 1e-6 and rel(dlv, lr.grad) < 1e-6
    for target in (0.9, 0.1):
        p = torch.rand(64, device="cuda") * 0.98 + 0.01
        p[0], p[1] = 1.0, 0.0  # saturated outputs: log clamped at -100 (nn.BCELoss)
        pr = p.clone().requires_grad_(True)
        refl = F.binary_cross_entropy(pr, torch.full((64,), target, device="cuda"))
        refl.backward()
        loss.zero_()
        dp = torch.empty_like(p)
        stat = torch.zeros((), device="cuda")
        ops.bce_const(p, target, loss, dprob=dp, stat=stat)
        assert abs(float(loss) - float(refl)) < 1e-5 * abs(float(refl))
        assert rel(dp, pr.grad) < 1e-5 and abs(float(stat) - float(p.sum())) < 1e-3


def test_adam_matches_torch(ops):
    torch.manual_seed(0)
    n = 100003
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step, 1.0, shadow)
    assert rel(p, ref.detach()) < 1e-6
    assert torch.equal(shadow, p.bfloat16())
    st = opt.state[ref]
    assert rel(m, st["exp_avg"]) < 1e-6 and rel(v, st["exp_avg_sq"]) < 1e-6


def test_adam_segments_and_bf16_gradients(ops):
    """dm_adam_step_ex: one optimizer step applied segment by segment (device step counter incremented once), with a
    bf16 gradient on the middle segment -- equals torch.optim.Adam fed the same (bf16-rounded) gradient."""
    torch.manual_seed(1)
    n, a, b = 4096 + 40000 + 1003, 4096, 4096 + 40000
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    step_dev = torch.zeros((), dtype=torch.int32, device="cuda")
    for _ in range(4):
        g = torch.randn(n, device="cuda")
        g16 = g[a:b].bfloat16()
        gref = g.clone()
        gref[a:b] = g16.float()
        ref.grad = gref
        opt.step()
        for i, (lo, hi, gg) in enumerate(((0, a, g[0:a]), (a, b, g16), (b, n, g[b:n]))):
            ops.adam_step(p[lo:hi], gg, m[lo:hi], v[lo:hi], 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, shadow[lo:hi],
                          step_dev=step_dev, count_step=(i == 0))
    assert int(step_dev) == 4
    assert rel(p, ref.detach()) < 1e-6 and torch.equal(shadow, p.bfloat16())


@pytest.mark.parametrize("rows", [8, 64, 200])
def test_bn1d_on_column_blocks_with_pre_bias(ops, rows):
    """dm_bn1d_forward / dm_bn1d_backward: BatchNorm1d + ReLU of one 2048-column block of a wider fp32 matrix whose
    producing GEMM did not add the Linear bias (the encoder's two heads from one N = 4096 GEMM, model.py:460-471),
    against torch.nn.functional.batch_norm on (block + bias)."""
    import torch.nn.functional as F

    torch.manual_seed(rows)
    c, ld = 2048, 4096
    acc = torch.randn(rows, ld, device="cuda") * 2 + 0.5
    for head in range(2):
        col0 = head * c
        bias = torch.randn(c, device="cuda")
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
        rm, rv = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
        nbt = torch.zeros((), dtype=torch.int64, device="cuda")
        x = (acc[:, col0:col0 + c] + bias).clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        rm_ref, rv_ref = rm.clone(), rv.clone()
        ref = F.relu(F.batch_norm(x, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5))
        out, ss, mi = ops.bn1d_forward_cols(acc, col0, c, bias, gamma, beta, rm, rv, nbt, 1)
        assert float((out.float() - ref).norm() / ref.norm()) < 4e-3
        assert torch.allclose(rm, rm_ref, atol=1e-5, rtol=1e-4) and torch.allclose(rv, rv_ref, atol=1e-5, rtol=1e-4)
        assert int(nbt) == 1
        dout = torch.randn(rows, c, device="cuda").bfloat16()
        ref.backward(dout.float())
        dy = torch.full((rows, ld), 7.0, device="cuda", dtype=torch.bfloat16)
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        ops.bn1d_backward_cols(dout, acc, col0, c, ss, mi, 1, 0.2, dy, dg, db)
        got = dy[:, col0:col0 + c].float()
        assert float((got - x.grad).norm() / x.grad.norm()) < 6e-3
        other = dy[:, (1 - head) * c:(2 - head) * c]
        assert bool((other == 7.0).all())  # the other head's columns are untouched
        assert float((dg - gr.grad).norm() / gr.grad.norm()) < 2e-3 and float((db - br.grad).norm() / br.grad.norm()) < 2e-3


@pytest.mark.parametrize("rows", [5, 64, 256])
def test_linear_pair_kernels(ops, rows):
    """dm_linear_pair_forward / backward: Linear(2048, 128) of both encoder heads (model.py:464,470) in one launch against
    torch on the same bf16 operands (fp32 accumulation)."""
    torch.manual_seed(rows)
    n, k = 128, 2048
    xs = [torch.randn(rows, k, device="cuda").bfloat16() for _ in range(2)]
    ws = [(torch.randn(n, k, device="cuda") * 0.03).bfloat16() for _ in range(2)]
    bs = [torch.randn(n, device="cuda") for _ in range(2)]
    o0, o1 = ops.linear_pair_forward(xs[0], xs[1], ws[0], ws[1], bs[0], bs[1])
    for o, x, w, b in zip((o0, o1), xs, ws, bs):
        ref = x.float() @ w.float().t() + b
        assert float((o - ref).norm() / ref.norm()) < 1e-5
    ds = [torch.randn(rows, n, device="cuda") for _ in range(2)]
    dws = [torch.full((n, k), 0.5, device="cuda") for _ in range(2)]  # accumulated into
    dbs = [torch.full((n,), -1.0, device="cuda") for _ in range(2)]
    dx0, dx1 = ops.linear_pair_backward(ds[0], ds[1], xs[0], xs[1], ws[0], ws[1], dws[0], dws[1], dbs[0], dbs[1])
    for dx, d, x, w, dw, db in zip((dx0, dx1), ds, xs, ws, dws, dbs):
        ref_dx = d @ w.float()
        assert float((dx.float() - ref_dx).norm() / ref_dx.norm()) < 4e-3  # one bf16 rounding of the result
        ref_dw = d.t() @ x.float() + 0.5
        assert float((dw - ref_dw).norm() / ref_dw.norm()) < 1e-5
        assert torch.allclose(db, d.sum(0) - 1.0, atol=1e-4, rtol=1e-5)
    # input gradient only (the discarded-weight-gradient phases): dw / db untouched
    dx0b, _ = ops.linear_pair_backward(ds[0], ds[1], xs[0], xs[1], ws[0], ws[1])
    assert torch.equal(dx0b, dx0)

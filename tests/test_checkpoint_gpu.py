"""GPU: checkpoint interchange with the reference's format (new_betavaegan.py:203-209, 222-228):
a `model_N.tar`-style dict written by the (oracle) reference loop is loaded into the kernel-backed modules + fused
trainer and training continues from it; the trainer's state goes back into stock torch modules / torch.optim.Adam.
Also pins that the tap-major storage of the 5x5 conv weights inside the trainer's flat buffers is invisible from
outside: state_dict tensors have the reference shapes and values."""
import io

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_reference_checkpoint_roundtrip():
    from disentangle_mlp_b200 import model as dm
    from disentangle_mlp_b200 import trainer as tr
    from oracle import nets, steps

    b, opt = 8, steps.make_opt()
    x = steps.synthetic_batch(b, 1234)
    torch.manual_seed(999)
    rEG, rD = nets.VAE(opt), nets.Discriminator_celeba(opt)
    rEG.apply(nets.weights_init)
    rD.apply(nets.weights_init)
    oEG, oD = torch.optim.Adam(rEG.parameters(), lr=1e-3), torch.optim.Adam(rD.parameters(), lr=1e-3)

    def rands(s):
        g = torch.Generator().manual_seed(500 + s)
        return [torch.randn(b, 128, generator=g) for _ in range(3)]

    for s in range(2):  # the reference trains two steps, then writes its checkpoint (through a real torch.save)
        steps.betavaegan_step(rEG, rD, oEG, oD, x, 25.0, 0.9, 0.1, *rands(s))
    buf = io.BytesIO()
    torch.save({"epoch": 1, "encoder_decoder_model": rEG.state_dict(), "discriminator_model": rD.state_dict(),
                "encoder_decoder_optimizer": oEG.state_dict(), "discriminator_optimizer": oD.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)

    # ---- load it into the kernel-backed modules AFTER the trainer has re-homed their parameters
    mEG, mD = dm.VAE(opt).cuda(), dm.Discriminator_celeba(opt).cuda()
    T = tr.BetaVAEGANTrainer(mEG, mD, beta=25.0, lr=1e-3)
    mEG.load_state_dict(ck["encoder_decoder_model"])
    mD.load_state_dict(ck["discriminator_model"])
    T.feg.load_optimizer_state_dict(ck["encoder_decoder_optimizer"])
    T.fd.load_optimizer_state_dict(ck["discriminator_optimizer"])
    # (no explicit refresh: load_state_dict's post-hook re-derives the bf16 shadow and the operand packs)
    assert T.feg.step_count == 4 and T.fd.step_count == 2
    for (n, p), (_, q) in zip(mEG.state_dict().items(), rEG.state_dict().items()):
        assert p.shape == q.shape and torch.equal(p.cpu(), q), n
    w = dict(mEG.named_parameters())["deconv2.weight"]
    assert tuple(w.shape) == (256, 128, 5, 5) and not w.is_contiguous()  # tap-major storage, reference-shaped view

    # ---- both continue for one step from the checkpoint on the same data / noise / labels
    n, e1, e2 = rands(2)
    ref = steps.betavaegan_step(rEG, rD, oEG, oD, x, 25.0, 0.9, 0.1, n, e1, e2)
    got = {k: float(v) for k, v in T.step(x.cuda(), 0.9, 0.1, n.cuda(), e1.cuda(), e2.cuda()).items()}
    for k, tol in (("errD_real", 5e-3), ("errD_fake", 5e-3), ("recon_dec", 2e-2), ("recon_enc", 5e-2)):
        assert abs(got[k] - ref[k]) <= tol * abs(ref[k]), (k, got[k], ref[k])

    # ---- and back: the trainer's state into stock torch modules and torch.optim.Adam (through torch.save again)
    buf = io.BytesIO()
    torch.save({"encoder_decoder_model": mEG.state_dict(), "discriminator_model": mD.state_dict(),
                "encoder_decoder_optimizer": T.feg.optimizer_state_dict(),
                "discriminator_optimizer": T.fd.optimizer_state_dict()}, buf)
    buf.seek(0)
    ck2 = torch.load(buf, map_location="cpu", weights_only=False)
    fEG, fD = nets.VAE(opt), nets.Discriminator_celeba(opt)
    fEG.load_state_dict(ck2["encoder_decoder_model"])
    fD.load_state_dict(ck2["discriminator_model"])
    fo = torch.optim.Adam(fEG.parameters(), lr=1e-3)
    fo.load_state_dict(ck2["encoder_decoder_optimizer"])
    for (n_, p), (_, q) in zip(fEG.named_parameters(), mEG.named_parameters()):
        assert torch.equal(p.detach(), q.detach().cpu()), n_
    st = fo.state_dict()["state"]
    assert len(st) == 42 and int(st[0]["step"]) == 6
    i = [n_ for n_, _ in fEG.named_parameters()].index("deconv2.weight")
    assert tuple(st[i]["exp_avg"].shape) == (256, 128, 5, 5)
    # sampling / reconstruction under no_grad through the module API (utils/utils.py:13-32) on parameters
    # that live, re-homed and tap-major, inside the trainer's flat buffers
    code = torch.randn(8, 128, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        img, img_ref = mEG.decode(code.cuda()).cpu(), fEG.decode(code)
        rec, mu, logvar = mEG(x.cuda())  # draws its own eps (models/model.py:534): only shapes are comparable
    assert float((img - img_ref).norm() / img_ref.norm()) < 1.5e-2
    assert tuple(rec.shape) == (b, 3, 64, 64) and tuple(mu.shape) == (b, 128) and bool(torch.isfinite(rec).all())
    # the imported moments continue training in stock torch without error
    steps.betavaegan_step(fEG, fD, fo, torch.optim.Adam(fD.parameters(), lr=1e-3), x, 25.0, 0.9, 0.1, *rands(3))


@pytest.mark.parametrize("kind", ["gan", "vae", "betavaegan"])
def test_checkpoint_key_sets_and_dataparallel_prefix(kind, tmp_path):
    """The three scripts' `model_N.tar` formats, key for key, INCLUDING the "module." prefix the reference's
    nn.DataParallel wrappers put on some state_dicts (new_gan.py:169-174 both nets; new_betavaegan.py:222-228 netD
    only; new_vae.py:88-91 none): a reference-written file resumes here (--load_path) and a file written here loads
    into the reference's DataParallel-wrapped modules."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiments"))
    import _common as C
    from disentangle_mlp_b200 import model as dm
    from disentangle_mlp_b200 import trainer as tr
    from oracle import nets, steps

    b, opt = 8, steps.make_opt()
    x = steps.synthetic_batch(b, 77)
    torch.manual_seed(999)
    if kind == "gan":
        refs = (nets.Generator_celeba(opt), nets.Discriminator_celeba(opt))
        mine = (dm.Generator_celeba(opt).cuda(), dm.Discriminator_celeba(opt).cuda())
    elif kind == "vae":
        refs = (nets.VAE(opt),)
        mine = (dm.VAE(opt).cuda(),)
    else:
        refs = (nets.VAE(opt), nets.Discriminator_celeba(opt))
        mine = (dm.VAE(opt).cuda(), dm.Discriminator_celeba(opt).cuda())
    for r in refs:
        r.apply(nets.weights_init)
    ropts = [torch.optim.Adam(r.parameters(), lr=3e-4) for r in refs]
    g = torch.Generator().manual_seed(8)
    rands = [torch.randn(b, 128, generator=g) for _ in range(3)]
    # the reference trains one step and writes its checkpoint the way ITS script does (DataParallel prefixes)
    if kind == "gan":
        steps.gan_step(refs[0], refs[1], ropts[0], ropts[1], x, 0.9, 0.1, rands[0])
        T = tr.GANTrainer(*mine, lr=3e-4)
        fps = (T.fg, T.fd)
    elif kind == "vae":
        steps.vae_step(refs[0], ropts[0], x, rands[0])
        T = tr.VAETrainer(mine[0], lr=3e-4)
        fps = (T.fp,)
    else:
        steps.betavaegan_step(refs[0], refs[1], ropts[0], ropts[1], x, 25.0, 0.9, 0.1, *rands)
        T = tr.BetaVAEGANTrainer(*mine, beta=25.0, lr=3e-4)
        fps = (T.feg, T.fd)
    ck = {"epoch": 3}
    for (mk, prefixed, ok), r, o in zip(C.CKPT_KEYS[kind], refs, ropts):
        sd = torch.nn.DataParallel(r).state_dict() if prefixed else r.state_dict()
        assert all(k.startswith("module.") for k in sd) == prefixed
        ck[mk], ck[ok] = sd, o.state_dict()
    path = tmp_path / "model_3.tar"
    torch.save(ck, path)
    # ---- resume here from the reference's file
    assert C.load_checkpoint(kind, str(path), mine, fps, "cuda") == 3
    for m, r in zip(mine, refs):
        for (n, p), (_, q) in zip(m.state_dict().items(), r.state_dict().items()):
            assert torch.equal(p.cpu(), q.cpu()), n  # (DataParallel moved a prefixed reference module to the GPU)
    for fp, o in zip(fps, ropts):
        st = o.state_dict()["state"]
        assert fp.step_count == int(st[0]["step"])
    T.step(x.cuda()) if kind == "vae" else T.step(x.cuda(), 0.9, 0.1)  # training continues from it
    # ---- and a file written here has the reference's key set and loads into ITS (DataParallel-wrapped) modules
    out = C.save_checkpoint(kind, str(tmp_path / "out"), 4, mine, fps)
    ck2 = torch.load(out, map_location="cpu", weights_only=False)
    assert set(ck2) == set(ck)
    for (mk, prefixed, ok), r in zip(C.CKPT_KEYS[kind], refs):
        assert set(ck2[mk]) == set(ck[mk]), mk
        target = torch.nn.DataParallel(r) if prefixed else r
        target.load_state_dict(ck2[mk])  # strict: raises on any key mismatch
        torch.optim.Adam(r.parameters(), lr=3e-4).load_state_dict(ck2[ok])

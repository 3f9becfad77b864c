"""GPU probe: kernel-backed modules vs the CPU oracle on identical weights/inputs; prints relative errors."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from disentangle_mlp_b200 import model as dm
from oracle import nets, steps


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def compare_grads(mine, ref, tag):
    worst = 0.0
    gscale = max(float(p.grad.norm()) for p in ref.parameters())
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        if p.grad is None:
            print(f"  {tag}.{n}: NO GRAD")
            continue
        r = rel(p.grad, q.grad)
        a = float((p.grad.cpu() - q.grad).norm()) / gscale
        flag = "" if (r < 3e-2 or a < 1e-4) else "   <-- "
        print(f"  {tag}.{n:28s} rel {r:9.3e}  abs/maxnorm {a:9.3e} |ref| {float(q.grad.norm()):9.3e}{flag}")
        if a >= 1e-4:
            worst = max(worst, r)
    return worst


def main():
    b = int(os.environ.get("B", "8"))
    torch.manual_seed(999)
    opt = steps.make_opt()
    ref_vae, ref_d = nets.VAE(opt), nets.Discriminator_celeba(opt)
    ref_vae.apply(nets.weights_init)
    ref_d.apply(nets.weights_init)
    import copy
    ref_d2 = copy.deepcopy(ref_d)
    my_vae, my_d = dm.VAE(opt).cuda(), dm.Discriminator_celeba(opt).cuda()
    my_vae.load_state_dict(ref_vae.state_dict())
    my_d.load_state_dict(ref_d.state_dict())
    x = steps.synthetic_batch(b, 1234)
    eps = torch.randn(b, 128)

    # ---- discriminator
    prob_r, feat_r = ref_d(x)
    (prob_r.sum() + 0.01 * feat_r.pow(2).sum()).backward()
    xg = x.cuda().requires_grad_(True)
    prob, feat = my_d(xg)
    (prob.sum() + 0.01 * feat.pow(2).sum()).backward()
    print(f"D prob rel {rel(prob, prob_r):.3e} feat rel {rel(feat, feat_r):.3e}")
    xr = x.clone().requires_grad_(True)
    pr, fr = ref_d2(xr)
    (pr.sum() + 0.01 * fr.pow(2).sum()).backward()
    print(f"D dx rel {rel(xg.grad, xr.grad):.3e}")
    wd = compare_grads(my_d, ref_d, "D")
    for k, v in ref_d.state_dict().items():
        if "running" in k or "tracked" in k:
            r = rel(my_d.state_dict()[k], v)
            if r > 1e-2:
                print("  BUFFER MISMATCH", k, r)

    # ---- VAE with injected eps
    mu_r, lv_r = ref_vae.encode(x)
    rec_r = ref_vae.decode(mu_r + eps * torch.exp(0.5 * lv_r))
    loss_r = torch.nn.functional.mse_loss(rec_r, x, reduction="sum") + steps.kld_sum(mu_r, lv_r)
    loss_r.backward()
    xc = x.cuda()
    mu, lv = my_vae.encode(xc)
    z = dm._ReparamFn.apply(mu, lv, eps.cuda())
    rec = my_vae.decode(z)
    loss = torch.nn.functional.mse_loss(rec, xc, reduction="sum") + steps.kld_sum(mu, lv)
    loss.backward()
    print(f"VAE mu rel {rel(mu, mu_r):.3e} logvar rel {rel(lv, lv_r):.3e} recon rel {rel(rec, rec_r):.3e} "
          f"loss {float(loss):.4f} vs {float(loss_r):.4f}")
    wv = compare_grads(my_vae, ref_vae, "VAE")
    for k, v in ref_vae.state_dict().items():
        if "running" in k or "tracked" in k:
            r = rel(my_vae.state_dict()[k], v)
            if r > 1e-2:
                print("  BUFFER MISMATCH", k, r)
    print(f"worst significant grad rel: D {wd:.3e}  VAE {wv:.3e}")


if __name__ == "__main__":
    main()

"""GPU: TEACHER-FORCED 200-step check of the fused beta-VAE-GAN / GAN steps against the oracle.

The free-running 200-step curves (test_curves_gpu.py) are chaotic in their GAN-side quantities (KL, Dis_l, errD/errG):
an fp32 re-run of the oracle with inputs perturbed by 1e-3 already deviates by 6-25 % there, so those curves cannot
detect a kernel bug.  Here the oracle runs its own fp32 trajectory on the host, and BEFORE EVERY STEP the CUDA
trainer is reset to the oracle's exact state (parameters, BatchNorm buffers, Adam moments and step count); both then
take the same step on identical data, labels, noise and eps.  What is compared is therefore the ONE-STEP map at 200
different points of a real training trajectory: every loss of the step and the parameter update it produces.
North-star tolerance (bf16): 1e-2 on the losses."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _flat(module):
    return torch.cat([p.detach().flatten().cpu() for p in module.parameters()])


def _force(module, fp, ref_module, ref_opt):
    """CUDA module + fused optimizer state := the oracle's state"""
    module.load_state_dict(ref_module.state_dict())  # (post-hook refreshes the bf16 shadow / operand packs)
    sd = ref_opt.state_dict()
    if sd["state"]:
        fp.load_optimizer_state_dict(sd)


def _run(workload, nsteps, b):
    from disentangle_mlp_b200 import model as dm
    from disentangle_mlp_b200 import trainer as tr
    from oracle import nets, steps

    opt = steps.make_opt()
    x = steps.synthetic_batch(b, 1234)
    xg = x.cuda()
    torch.manual_seed(999)
    np.random.seed(999)
    if workload == "betavaegan":
        ra, rd = nets.VAE(opt), nets.Discriminator_celeba(opt)
        ma = dm.VAE(opt).cuda()
        lr = 1e-3
    else:
        ra, rd = nets.Generator_celeba(opt), nets.Discriminator_celeba(opt)
        ma = dm.Generator_celeba(opt).cuda()
        lr = 3e-4
    ra.apply(nets.weights_init)
    rd.apply(nets.weights_init)
    md = dm.Discriminator_celeba(opt).cuda()
    oa, od = torch.optim.Adam(ra.parameters(), lr=lr), torch.optim.Adam(rd.parameters(), lr=lr)
    if workload == "betavaegan":
        T = tr.BetaVAEGANTrainer(ma, md, beta=25.0, lr=lr)
        fa, fd = T.feg, T.fd
    else:
        T = tr.GANTrainer(ma, md, lr=lr)
        fa, fd = T.fg, T.fd
    devs, upd = {}, {"a": [], "d": []}
    for s in range(nsteps):
        _force(ma, fa, ra, oa)
        _force(md, fd, rd, od)
        pa0, pd0 = _flat(ra), _flat(rd)
        real, fake = steps.draw_labels()
        g = torch.Generator().manual_seed(10_000 + s)
        rands = [torch.randn(b, 128, generator=g) for _ in range(T.n_rands)]
        if workload == "betavaegan":
            r = steps.betavaegan_step(ra, rd, oa, od, x, 25.0, real, fake, *rands)
        else:
            r = steps.gan_step(ra, rd, oa, od, x, real, fake, *rands)
        m = {k: float(v) for k, v in T.step(xg, real, fake, *[t.cuda() for t in rands]).items()}
        for k in r:
            devs.setdefault(k, []).append(abs(m[k] - r[k]) / (abs(r[k]) + 1e-9))
        T.sync()  # (apply the deferred big-tensor update before the parameters are read)
        for key, mm, rr, p0 in (("a", ma, ra, pa0), ("d", md, rd, pd0)):
            du_ref = _flat(rr) - p0
            du = _flat(mm) - p0
            upd[key].append(float((du - du_ref).norm() / (du_ref.norm() + 1e-30)))
    return devs, upd


@pytest.mark.parametrize("workload,nsteps", [("betavaegan", 200), ("gan", 100)])
def test_teacher_forced_one_step_map(workload, nsteps):
    devs, upd = _run(workload, nsteps, 16)
    report = {k: {"median": float(np.median(v)), "p90": float(np.percentile(v, 90)), "max": float(np.max(v))}
              for k, v in devs.items()}
    report["update_rel_err"] = {k: {"median": float(np.median(v)), "max": float(np.max(v)),
                                    "median_after_step20": float(np.median(v[20:]))} for k, v in upd.items()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/teacher_forced_{workload}.json", "w") as f:
        json.dump(report, f, indent=1)
    print(workload, json.dumps(report))
    # losses of the one-step map: north-star bf16 tolerance in the median over the trajectory, and a bound on the
    # 90th percentile (single steps where D sits at the BCE clamp or a loss crosses ~0 have large RELATIVE error)
    # (measured on B200, round 2: medians <= 1e-3 for every quantity incl. KL 9e-5 and Dis_l 9e-4; worst single step 2.4e-2)
    for k, v in devs.items():
        assert np.all(np.isfinite(v)), k
        assert np.median(v) <= 1e-2, (k, report[k])
        assert np.percentile(v, 90) <= 1e-2, (k, report[k])
        assert np.max(v) <= 1e-1, (k, report[k])
    # parameter update of one step (three Adam updates): Adam normalises every element's step to ~lr, so elements
    # whose gradient is at the bf16 noise level move in a noise-determined direction in the reference too
    for k, v in upd.items():  # (measured medians: EG 0.03, D 1e-5)
        assert np.median(v) <= 0.15, (k, report["update_rel_err"][k])

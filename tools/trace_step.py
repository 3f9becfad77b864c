"""Kernel timeline of the CUDA-graph step (torch.profiler / CUPTI): per-kernel totals and the idle gaps between kernels.
    python tools/trace_step.py [workload] [batch] [graph:0|1]"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "betavaegan"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    graph = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    torch.manual_seed(999)
    np.random.seed(999)
    opt = dm.default_opt()
    eg, d = dm.VAE(opt), dm.Discriminator_celeba(opt)
    eg.apply(dm.weights_init)
    d.apply(dm.weights_init)
    T = tr.BetaVAEGANTrainer(eg.cuda(), d.cuda(), beta=1.0, lr=1e-3)
    if graph:
        T.enable_graph(b)
    x = (torch.rand(b, 3, 64, 64) * 2 - 1).cuda()
    for _ in range(5):
        T.step(x)
    torch.cuda.synchronize()
    nsteps = 3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(nsteps):
            T.step(x)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no CUDA events captured")
        return
    tot = defaultdict(lambda: [0, 0.0])
    busy = 0.0
    for s, e, n in ks:
        tot[n.split("(")[0][:70]][0] += 1
        tot[n.split("(")[0][:70]][1] += e - s
        busy += e - s
    span = ks[-1][1] - ks[0][0]
    gaps = [ks[i + 1][0] - ks[i][1] for i in range(len(ks) - 1)]
    gaps_pos = [g for g in gaps if g > 0]
    print(f"{len(ks)} kernels over {nsteps} steps; span {span / nsteps / 1e3:.3f} ms/step, busy {busy / nsteps / 1e3:.3f} ms/step, "
          f"gaps {sum(gaps_pos) / nsteps / 1e3:.3f} ms/step (median gap {np.median(gaps):.2f} us)")
    if os.environ.get("TRACE_GAPS"):
        big = sorted(((ks[i + 1][0] - ks[i][1], i) for i in range(len(ks) - 1)), reverse=True)[:int(os.environ["TRACE_GAPS"])]
        for g, i in big:
            print(f"gap {g:8.1f} us after [{ks[i][2][:50]}] ({ks[i][1] - ks[i][0]:.1f} us) before [{ks[i + 1][2][:50]}]")
    dump = os.environ.get("TRACE_DUMP")  # path: ordered list "start_us dur_us name" of ONE step's kernels
    if dump:
        per_step = len(ks) // nsteps
        t0 = ks[0][0]
        with open(dump, "w") as f:
            for s_, e, n in ks[:per_step]:
                f.write(f"{s_ - t0:10.1f} {e - s_:8.1f} {n[:90]}\n")
    show = os.environ.get("TRACE_SHOW")  # substring: list that kernel's individual launches (one step) in order
    if show:
        per = [(e - s_) for s_, e, n in ks if show in n]
        per = per[: len(per) // nsteps]
        print(f"{show}: " + " ".join(f"{d:.1f}" for d in per))
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / nsteps / 1e3:8.3f} ms/step  x{c / nsteps:6.1f}  avg {t / c:7.1f} us  {n}")


if __name__ == "__main__":
    main()

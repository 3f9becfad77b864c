"""2-rank data-parallel step under torch.profiler: how much of the NCCL all-reduce time overlaps compute kernels.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/trace_dp.py [batch] [graph]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from disentangle_mlp_b200 import model as dm
from disentangle_mlp_b200 import trainer as tr


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    graph = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
    dist.barrier()
    nsteps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(nsteps):
            T.step(x)
        torch.cuda.synchronize()
    if dist.get_rank() == 0:
        evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events()
               if e.device_type == torch.autograd.DeviceType.CUDA]
        nccl = [(s, e) for s, e, n in evs if "nccl" in n.lower()]
        comp = sorted((s, e) for s, e, n in evs if "nccl" not in n.lower())
        span = max(e for _, e, _ in evs) - min(s for s, _, _ in evs)
        tn = sum(e - s for s, e in nccl)
        ov = 0.0
        for s, e in nccl:
            for cs, ce in comp:
                if ce <= s:
                    continue
                if cs >= e:
                    break
                ov += min(e, ce) - max(s, cs)
        print(f"span {span / nsteps / 1e3:.3f} ms/step; nccl kernels {len(nccl) / nsteps:.1f}/step, {tn / nsteps / 1e3:.3f} ms/step, "
              f"of which overlapped with compute kernels {ov / nsteps / 1e3:.3f} ms/step; compute busy "
              f"{sum(e - s for s, e in comp) / nsteps / 1e3:.3f} ms/step")
        for s, e in nccl[: len(nccl) // nsteps]:
            print(f"   nccl kernel {(e - s):8.1f} us")
        # idle time of the compute side: gaps between consecutive compute kernels (union of intervals), largest first
        merged = []
        for cs, ce in comp:
            if merged and cs <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], ce)
            else:
                merged.append([cs, ce])
        gaps = sorted(((merged[i + 1][0] - merged[i][1], merged[i][1]) for i in range(len(merged) - 1)), reverse=True)
        print(f"compute idle (no non-NCCL kernel running) {sum(g for g, _ in gaps) / nsteps / 1e3:.3f} ms/step; largest gaps:")
        by_end = sorted(evs, key=lambda t: t[1])
        for g, at in gaps[:24]:
            prev = [n for s_, e_, n in evs if abs(e_ - at) < 0.01 and "nccl" not in n.lower()]
            nxt = [n for s_, e_, n in evs if abs(s_ - (at + g)) < 0.01 and "nccl" not in n.lower()]
            during = [f"{n[:28]}({e_ - s_:.0f}us)" for s_, e_, n in evs if "nccl" in n.lower() and s_ < at + g and e_ > at]
            print(f"   {g:7.1f} us after [{(prev or ['?'])[0][:40]}] before [{(nxt or ['?'])[0][:40]}] nccl: {during}")
        dump = os.environ.get("TRACE_DUMP")
        if dump:
            ks = sorted(evs)
            t0 = ks[0][0]
            with open(dump, "w") as f:
                for s_, e_, n in ks[: len(ks) // nsteps]:
                    f.write(f"{s_ - t0:10.1f} {e_ - s_:8.1f} {n[:90]}\n")
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()

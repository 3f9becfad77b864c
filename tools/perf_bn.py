"""Micro-benchmark of the BatchNorm streaming kernels on the step's layer shapes: achieved GB/s (algorithmic bytes)
per kernel with an L2 flush between repetitions (cold) and back-to-back (L2-warm where the tensor fits).
    DM_BN_STREAM_BLOCKS=4 python tools/perf_bn.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from disentangle_mlp_b200 import ops

SHAPES = [(192 * 4096, 32, 3), (128 * 4096, 32, 2), (64 * 4096, 32, 1), (192 * 1024, 128, 3), (64 * 1024, 128, 1),
          (192 * 256, 256, 3), (64 * 256, 256, 1), (64 * 64, 256, 1)]


def timeit(fn, flush, reps=10):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    print(f"DM_BN_STREAM_BLOCKS={os.environ.get('DM_BN_STREAM_BLOCKS', 'default')}")
    print(f"{'rows x c (groups)':>24s} | {'stats':>14s} | {'apply':>14s} | {'bwd reduce+apply':>18s}   (us cold / us warm, GB/s cold)")
    for rows, c, g in SHAPES:
        rg = rows // g
        y = torch.randn(rows, c, device=dev).bfloat16()
        dout = torch.randn(rows, c, device=dev).bfloat16()
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        sc = ops.bn_scratch(c, g, dev)
        out = torch.empty_like(y)
        res = {}

        site = ops.BnSite(sc, g, rg, c, gamma, beta, rm, rv, None)

        def stats():
            ops.bn_stats(y, c, site)

        def apply():
            ops.bn_apply_act(y, rg, c, site.scale_shift, 2, 0.2, out=out, groups=g)

        def both():
            stats()
            apply()

        both()
        ss, mi = site.scale_shift, site.mean_invstd

        def bwd():
            ops.bn_backward(dout, y, rg, c, ss, mi, 2, 0.2, None, None, out=out, groups=g, scratch=sc)

        n = rows * c
        t_both_c, t_both_w = timeit(both, flush), timeit(both, None)
        t_st_c = timeit(stats, flush)
        t_st_w = timeit(stats, None)
        t_bwd_c, t_bwd_w = timeit(bwd, flush), timeit(bwd, None)
        ap_c, ap_w = t_both_c - t_st_c, t_both_w - t_st_w
        print(f"{rows:>10d} x {c:<4d} ({g})      | {t_st_c:6.1f}/{t_st_w:6.1f} {2 * n / t_st_c / 1e3:5.0f} | "
              f"{ap_c:6.1f}/{ap_w:6.1f} {4 * n / max(ap_c, 1e-3) / 1e3:5.0f} | {t_bwd_c:6.1f}/{t_bwd_w:6.1f} {10 * n / t_bwd_c / 1e3:6.0f}")


if __name__ == "__main__":
    main()

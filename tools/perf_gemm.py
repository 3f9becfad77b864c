"""GPU micro-benchmark of the tap-GEMM kernel on the training step's main layer shapes (CUDA events, warm caches):
    python tools/perf_gemm.py [batch] ["ENV=V ENV2=V" ...]
Each env set is applied in-process (the library reads its DM_* knobs at every launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from disentangle_mlp_b200 import ops


def timeit(fn, reps=20):
    for _ in range(min(3, reps)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def cases(b):
    dev = "cuda"
    out = []
    m, n, k = 8192, 4096, 4096
    a = torch.randn(m, k, device=dev).bfloat16()
    w = torch.randn(n, k, device=dev).bfloat16()
    d = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    out.append((f"gemm_nt {m}x{n}x{k}", 2.0 * m * n * k, lambda: ops.gemm(ops.GEMM_NT, a, w, m, n, k, out=d)))
    for (tm, tn, tk) in ((b, 128, 2048), (b, 16384, 128)):
        ta = torch.randn(tm, tk, device=dev).bfloat16()
        tw = torch.randn(tn, tk, device=dev).bfloat16()
        td = torch.empty(tm, tn, device=dev, dtype=torch.bfloat16)
        out.append((f"gemm_nt {tm}x{tn}x{tk}", 2.0 * tm * tn * tk,
                    lambda ta=ta, tw=tw, td=td, tm=tm, tn=tn, tk=tk: ops.gemm(ops.GEMM_NT, ta, tw, tm, tn, tk, out=td)))
    for (hs, cs, cb, stride) in ((16, 256, 128, 2), (8, 256, 256, 2), (32, 128, 32, 2)):
        wt = torch.randn(cs, cb, 5, 5, device=dev) * 0.05
        small = torch.randn(b, hs, hs, cs, device=dev).bfloat16()
        big = torch.randn(b, hs * stride, hs * stride, cb, device=dev).bfloat16()
        w_down, w_up, _ = ops.pack_conv_weights(wt, cs, cb)
        g = ops.geom(b, hs, hs, cs, cb, stride)
        fl = 50.0 * b * hs * hs * cs * cb
        o_s, o_b = torch.empty_like(small), torch.empty_like(big)
        dw = torch.zeros(cs, cb, 5, 5, device=dev)
        tag = f"{cb}->{cs} {hs * stride}->{hs}"
        out.append((f"down  {tag}", fl, lambda g=g, big=big, w_down=w_down, o_s=o_s: ops.conv_down(g, big, w_down, None, out=o_s)))
        if cb == 32 and stride == 2:
            w_pair = ops.pack_down_pairs(w_down, cs, cb)
            out.append((f"downP {tag}", fl, lambda g=g, big=big, w_pair=w_pair, o_s=o_s: ops.conv_down(g, big, w_pair, None, out=o_s)))
        out.append((f"up    {tag}", fl, lambda g=g, small=small, w_up=w_up, o_b=o_b: ops.conv_up(g, small, w_up, None, out=o_b)))
        out.append((f"wgrad {tag}", fl, lambda g=g, small=small, big=big, dw=dw: ops.conv_wgrad(g, small, big, dw)))
        pk = torch.zeros(25, cs, cb, device=dev)
        out.append((f"wgradP {tag}", fl, lambda g=g, small=small, big=big, pk=pk: ops.conv_wgrad_packed(g, small, big, pk)))
    # the 16384 <-> 2048 Linear layers (weight-bandwidth bound): bytes-equivalent "TFLOP/s" column is not meaningful,
    # read the microseconds: W is 67 MB in bf16 (10 us at 6.5 TB/s), dW 134 MB in fp32 (21 us)
    from disentangle_mlp_b200 import engine
    K, N = 16384, 2048
    x = torch.randn(b, K, device=dev).bfloat16()
    dy = torch.randn(b, N, device=dev).bfloat16()
    wl = (torch.randn(N, K, device=dev) * 0.01).bfloat16()
    bias = torch.zeros(N, device=dev)
    dwl = torch.empty(N, K, device=dev)
    dx = torch.empty(b, K, device=dev, dtype=torch.bfloat16)
    out.append((f"lin_fwd  {b}x{N}x{K}", 2.0 * b * N * K, lambda: engine.linear_forward(x, wl, bias, b, N, K)))
    out.append((f"lin_dgrad {b}x{K}x{N}", 2.0 * b * N * K, lambda: engine.linear_dgrad(dy, wl, b, N, K, out_dtype=torch.bfloat16)))
    out.append((f"lin_wgrad {K}x{N}x{b}", 2.0 * b * N * K, lambda: engine.linear_wgrad(dy, x, b, N, K, dwl, overwrite=True)))
    return out


def main():
    only = os.environ.get("PERF_ONLY")     # substring filter on the case name
    reps = int(os.environ.get("PERF_REPS", "20"))
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    envsets = sys.argv[2:] or [""]
    cs = [c for c in cases(b) if not only or only in c[0]]
    print(f"batch {b}")
    for es in envsets:
        added = []
        for kv in es.split():
            k, v = kv.split("=")
            os.environ[k] = v
            added.append(k)
        print(f"--- env [{es}]")
        for name, fl, fn in cs:
            us = timeit(fn, reps)
            grid, smem, stages = ops.last_plan()
            print(f"  {name:28s} {us:9.1f} us  {fl / us * 1e-6:8.1f} TFLOP/s   tiles {grid} stages {stages} smem {smem}")
        for k in added:
            del os.environ[k]


if __name__ == "__main__":
    main()

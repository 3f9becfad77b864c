"""GPU probe for the tap-GEMM kernel: runs each GEMM-class case in its own subprocess (so a hang or a
fault in one case cannot take the others down) and prints the relative error against torch.

    python tools/probe_gemm.py            # driver: all cases
    python tools/probe_gemm.py CASE       # one case in this process
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _fp32_refs():
    """torch references on CUDA must be true fp32: cudnn.allow_tf32 defaults to True (a ~1e-3 reference)."""
    import torch

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel_err(got, ref):
    import torch

    got = got.float()
    ref = ref.float()
    return float((got - ref).norm() / (ref.norm() + 1e-30)), float((got - ref).abs().max())


def case_gemm(layout, m, n, k, splits=1, out_bf16=False, bias=False, n_store=0, m_store=0):
    import torch

    from disentangle_mlp_b200 import ops

    _fp32_refs()
    torch.manual_seed(0)
    dev = "cuda"
    if layout == "nt":
        a = torch.randn(m, k, device=dev).bfloat16()
        b = torch.randn(n, k, device=dev).bfloat16()
        ref = a.float() @ b.float().t()
        lay = ops.GEMM_NT
    elif layout == "nn":
        a = torch.randn(m, k, device=dev).bfloat16()
        b = torch.randn(k, n, device=dev).bfloat16()
        ref = a.float() @ b.float()
        lay = ops.GEMM_NN
    else:
        a = torch.randn(k, m, device=dev).bfloat16()
        b = torch.randn(k, n, device=dev).bfloat16()
        ref = a.float().t() @ b.float()
        lay = ops.GEMM_TN
    bias_t = torch.randn(n, device=dev) if bias else None
    if bias:
        ref = ref + bias_t
    acc = splits > 1
    out = ops.gemm(lay, a, b, m, n, k, out_dtype=torch.bfloat16 if out_bf16 else torch.float32, accumulate=acc,
                   bias=bias_t, splits=splits, m_store=m_store, n_store=n_store)
    torch.cuda.synchronize()
    ref = ref[: (m_store or m), : (n_store or n)]
    return rel_err(out, ref)


def conv_refs(batch, hs, ws, cs, cb, stride, seed=0):
    import torch

    _fp32_refs()
    torch.manual_seed(seed)
    dev = "cuda"
    w = (torch.randn(cs, cb, 5, 5, device=dev) * 0.05)
    small = torch.randn(batch, hs, ws, cs, device=dev).bfloat16()
    big = torch.randn(batch, hs * stride, ws * stride, cb, device=dev).bfloat16()
    return w, small, big


def case_conv_down(batch, hs, ws, cs, cb, stride, bias=True):
    import torch
    import torch.nn.functional as F

    from disentangle_mlp_b200 import ops

    w, _, big = conv_refs(batch, hs, ws, cs, cb, stride)
    b = torch.randn(cs, device="cuda") if bias else None
    w_down, _, _ = ops.pack_conv_weights(w, cs, cb, want_up=False)
    g = ops.geom(batch, hs, ws, cs, cb, stride)
    out = ops.conv_down(g, big, w_down, b)
    torch.cuda.synchronize()
    ref = F.conv2d(big.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, stride=stride, padding=2)
    return rel_err(out, ref.permute(0, 2, 3, 1))


def case_conv_up(batch, hs, ws, cs, cb, stride, bias=True, out_f32=False):
    import torch
    import torch.nn.functional as F

    from disentangle_mlp_b200 import ops

    w, small, _ = conv_refs(batch, hs, ws, cs, cb, stride)
    b = torch.randn(cb, device="cuda") if bias else None
    _, w_up, _ = ops.pack_conv_weights(w, cs, cb, want_down=False)
    g = ops.geom(batch, hs, ws, cs, cb, stride)
    out = ops.conv_up(g, small, w_up, b, out_f32=out_f32)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(small.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, stride=stride, padding=2,
                             output_padding=stride - 1)
    return rel_err(out, ref.permute(0, 2, 3, 1))


def case_conv_wgrad(batch, hs, ws, cs, cb, stride, direct=True):
    import torch

    from disentangle_mlp_b200 import ops

    w, small, big = conv_refs(batch, hs, ws, cs, cb, stride)
    g = ops.geom(batch, hs, ws, cs, cb, stride)
    dw = torch.zeros(cs, cb, 5, 5, device="cuda")
    ops.conv_wgrad(g, small, big, dw, direct=direct)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(big.float().permute(0, 3, 1, 2), (cs, cb, 5, 5),
                                      small.float().permute(0, 3, 1, 2), stride=stride, padding=2)
    return rel_err(dw, ref)


CASES = {
    # plain GEMMs, K-major SW128
    "nt_128x128x64": lambda: case_gemm("nt", 128, 128, 64),
    "nt_128x128x256": lambda: case_gemm("nt", 128, 128, 256),
    "nt_256x512x1024_bias": lambda: case_gemm("nt", 256, 512, 1024, bias=True),
    "nt_bf16out": lambda: case_gemm("nt", 256, 256, 512, out_bf16=True, bias=True),
    "nt_ragged_m100_n48": lambda: case_gemm("nt", 100, 48, 128),
    "nt_splitk8": lambda: case_gemm("nt", 64, 2048, 16384, splits=8, bias=True),
    "nt_k32_sw64": lambda: case_gemm("nt", 256, 128, 32),
    "nt_k128_n32": lambda: case_gemm("nt", 4096, 32, 128, out_bf16=True, bias=True),
    # MN-major B
    "nn_128x128x64": lambda: case_gemm("nn", 128, 128, 64),
    "nn_64x16384x2048": lambda: case_gemm("nn", 64, 16384, 2048, out_bf16=True),
    "nn_splitk": lambda: case_gemm("nn", 64, 128, 16384, splits=32),
    # MN-major A and B
    "tn_128x128x64": lambda: case_gemm("tn", 128, 128, 64),
    "tn_128x256x128": lambda: case_gemm("tn", 128, 256, 128),
    "tn_2048x1024x64": lambda: case_gemm("tn", 2048, 1024, 64),
    "tn_acc_splitk": lambda: case_gemm("tn", 256, 256, 1024, splits=4),
    "tn_n32_oob_box": lambda: case_gemm("tn", 128, 32, 4096, splits=8, m_store=75),
    "tn_col": lambda: case_gemm("tn", 128, 64, 8192, splits=16, m_store=75, n_store=32),
    # convolutions
    "down_s1_c64": lambda: case_conv_down(2, 16, 16, 128, 64, 1),
    "down_s2_c64": lambda: case_conv_down(2, 16, 16, 128, 64, 2),
    "down_s2_c128_8x8": lambda: case_conv_down(4, 8, 8, 256, 128, 2),
    "down_s2_c32": lambda: case_conv_down(2, 32, 32, 128, 32, 2),
    "down_s2_c256_odd_batch": lambda: case_conv_down(3, 8, 8, 256, 256, 2),
    "up_s1_c64": lambda: case_conv_up(2, 16, 16, 64, 64, 1),
    "up_s2_c128": lambda: case_conv_up(2, 8, 8, 256, 128, 2),
    "up_s2_c32": lambda: case_conv_up(2, 32, 32, 128, 32, 2),
    "up_s1_32to3_f32": lambda: case_conv_up(2, 64, 64, 32, 3, 1, out_f32=True),
    "wgrad_s1": lambda: case_conv_wgrad(2, 16, 16, 128, 64, 1),
    "wgrad_s2": lambda: case_conv_wgrad(2, 16, 16, 256, 128, 2),
    "wgrad_s2_8x8": lambda: case_conv_wgrad(4, 8, 8, 256, 256, 2),
    "wgrad_s2_c32_pair": lambda: case_conv_wgrad(2, 32, 32, 128, 32, 2),
    "wgrad_s2_cb64": lambda: case_conv_wgrad(2, 16, 16, 128, 64, 2),
    "wgrad_packed_s2": lambda: case_conv_wgrad(2, 16, 16, 256, 128, 2, direct=False),
    "wgrad_packed_pair": lambda: case_conv_wgrad(2, 32, 32, 128, 32, 2, direct=False),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        name = sys.argv[1]
        t0 = time.time()
        rel, mx = CASES[name]()
        print(json.dumps({"case": name, "rel": rel, "max_abs": mx, "sec": round(time.time() - t0, 2)}))
        return
    results = []
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=90)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            if r.returncode == 0 and line.startswith("{"):
                d = json.loads(line)
                d["ok"] = d["rel"] < 2e-2
            else:
                d = {"case": name, "ok": False, "rc": r.returncode, "err": (r.stderr or r.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            d = {"case": name, "ok": False, "err": "TIMEOUT (hang)"}
        results.append(d)
        print(json.dumps(d), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_gemm.json", "w") as f:
        json.dump(results, f, indent=1)
    print("PASS" if all(d["ok"] for d in results) else "SOME FAILED")


if __name__ == "__main__":
    main()

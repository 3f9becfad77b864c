"""In-kernel phase timeline of the tap-GEMM kernel (debug build with -DDM_STAMPS, %globaltimer stamps per CTA):
    python tools/stamp_gemm.py [batch]
Build the stamped library first (CPU container):  tools/build_stamps.sh -> disentangle_mlp_b200/lib/libdm_b200_stamps.so
Columns are microseconds after the first CTA's entry: median / max over the CTAs of the launch."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DM_B200_LIB"] = os.path.join(ROOT, "disentangle_mlp_b200", "lib", "libdm_b200_stamps.so")
import numpy as np
import torch

from disentangle_mlp_b200 import _lib, ops

sys.path.insert(0, os.path.join(ROOT, "tools"))
import perf_gemm

NAMES = ["entry", "setup", "pdl", "tma0", "full0", "mma_end", "epi0", "epiL", "epi_end", "stat_end", "st_done", "prod_end",
         "exit"]


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    lib = _lib.load()
    lib.dm_debug_set_stamps.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
    lib.dm_debug_set_stamps(buf.data_ptr())
    only = os.environ.get("PERF_ONLY")
    print("us after first entry, median/max over CTAs   " + "  ".join(NAMES[1:]))
    for name, fl, fn in perf_gemm.cases(b):
        if only and only not in name:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        us = perf_gemm.timeit(fn, 10)
        buf.zero_()
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        grid, smem, stages = ops.last_plan()
        st = buf.cpu().numpy().reshape(-1, 16).astype(np.float64)
        st = st[st[:, 0] > 0]
        t0 = st[:, 0].min()
        cols = []
        for i in range(13):
            v = st[:, i]
            v = v[v > 0]
            if len(v) == 0:
                cols.append("   -   ")
            else:
                cols.append(f"{(np.median(v) - t0) / 1e3:5.1f}/{(v.max() - t0) / 1e3:5.1f}")
        extra = ""
        if (st[:, 14] > 0).any():
            extra = f" | last tile epilogue: {np.median(st[:, 14][st[:, 14] > 0]):.0f} clk, of which tcgen05.ld+wait {np.median(st[:, 13][st[:, 14] > 0]):.0f} clk, staging-buffer wait {np.median(st[:, 15][st[:, 14] > 0] // 100000):.0f}, pack + st.shared {np.median(st[:, 15][st[:, 14] > 0] % 100000):.0f}"
        print(f"{name:28s} {us:6.1f} us  ctas {len(st):4d} entry {cols[0]} | " + " ".join(cols[1:]) + extra)


if __name__ == "__main__":
    main()

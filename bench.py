#!/usr/bin/env python
"""Benchmark of the VAE-GAN / beta-VAE-GAN training step (BASELINE.json metric: train img/s per step, 64x64).

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's CUDA path (default)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                               # the reference's CPU step (oracle port)

N=1 workload = BASELINE.json configs[1]: VAE-GAN baseline (beta=1, Dis_l feature loss), 64x64, batch 64 on one
B200.  N>1: the same per-GPU batch on every rank (weak scaling; configs[2] is 64/GPU x 8), gradients
SUM-allreduced over NCCL before each of the three Adam steps.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMG = {"betavaegan": 20.01, "gan": 9.95, "vae": 3.64}  # algorithmic minimum, SURVEY.md §8d


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:  # nvidia-smi takes a moment to print its first sample
                time.sleep(0.02)
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class _Autocast:
    """bf16-autocast view of an oracle module for the stock-torch GPU baseline: forward / encode / decode run under
    torch.autocast(bfloat16) and return fp32 (nn.BCELoss refuses bf16 inputs).  (channels_last is not an option for
    the reference architecture: its `.view(-1, 16384)` flattens need NCHW-contiguous activations.)"""

    def __init__(self, module):
        self.m = module

    def _call(self, fn, x):
        import torch

        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = fn(x)
        return tuple(t.float() for t in o) if isinstance(o, tuple) else o.float()

    def __call__(self, x):
        return self._call(self.m, x)

    def encode(self, x):
        return self._call(self.m.encode, x)

    def decode(self, z):
        return self._call(self.m.decode, z)

    def zero_grad(self):
        self.m.zero_grad()


def build_oracle_step(workload, batch, threads, device="cpu", autocast=False, beta=1.0):
    """The reference's training step (oracle port: stock torch.nn modules + torch.optim.Adam, fp32): on the host
    cores (device "cpu": the CPU baseline / --impl reference), or on the GPU (the stock-PyTorch-on-B200 baseline)."""
    import numpy as np
    import torch

    from oracle import nets, steps

    if device == "cpu":
        torch.set_num_threads(threads)
    torch.manual_seed(999)
    np.random.seed(999)
    opt = steps.make_opt()
    x = steps.synthetic_batch(batch, 1234).to(device)
    wrap = _Autocast if autocast else (lambda m: m)
    if workload == "betavaegan":
        eg, d = nets.VAE(opt).to(device), nets.Discriminator_celeba(opt).to(device)
        eg.apply(nets.weights_init)
        d.apply(nets.weights_init)
        oeg, od = torch.optim.Adam(eg.parameters(), lr=1e-3), torch.optim.Adam(d.parameters(), lr=1e-3)
        weg, wd = wrap(eg), wrap(d)

        def step():
            real, fake = steps.draw_labels()
            return steps.betavaegan_step(weg, wd, oeg, od, x, beta, real, fake)
    elif workload == "gan":
        g, d = nets.Generator_celeba(opt).to(device), nets.Discriminator_celeba(opt).to(device)
        g.apply(nets.weights_init)
        d.apply(nets.weights_init)
        og, od = torch.optim.Adam(g.parameters(), lr=3e-4), torch.optim.Adam(d.parameters(), lr=3e-4)
        wg, wd = wrap(g), wrap(d)

        def step():
            real, fake = steps.draw_labels()
            return steps.gan_step(wg, wd, og, od, x, real, fake)
    else:
        m = nets.VAE(opt).to(device)
        m.apply(nets.weights_init)
        o = torch.optim.Adam(m.parameters(), lr=3e-4)
        wm = wrap(m)

        def step():
            return steps.vae_step(wm, o, x)
    return step


def time_torch_gpu(workload, batch, beta, warmup=3, steps_n=10):
    """The 'kernel to beat' of SURVEY.md §2b / BASELINE.md §4: the SAME step on stock torch.nn (cuDNN / cuBLAS)
    kernels on this B200 -- the oracle's modules moved to the GPU, nothing of this repo on the path -- in three
    precision modes.  CUDA-event timed; includes the reference loop's own `.item()` host syncs, as the reference
    would pay them.  Returns {mode: {"value": img/s, "ms_per_step": ...}}."""
    import torch

    out = {}
    modes = (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True))
    for name, tf32, ac in modes:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        try:
            step = build_oracle_step(workload, batch, 0, device="cuda", autocast=ac, beta=beta)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps_n):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps_n
            out[name] = {"value": round(batch / (ms / 1e3), 1), "unit": "img/s", "ms_per_step": round(ms, 3)}
        except Exception as e:  # noqa: BLE001 - a baseline that cannot run is reported, not fatal
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        finally:
            del step
            torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = False
    out["what"] = ("oracle/steps.py + oracle/nets.py (stock torch.nn / torch.optim.Adam restatement of the reference loop) "
                   f"on cuda:0, batch {batch}, {warmup} warm-up + {steps_n} timed steps, CUDA events")
    return out


def time_cpu(workload, batch, threads, warmup, steps_n, beta=1.0, budget_s=None):
    """Median step time of the CPU oracle step at `batch`.  budget_s: stop timing early (after >= 3 steps) when the
    wall clock passes it -- the sample gets shorter, the batch never smaller.  Returns (img/s, median s, steps timed)."""
    step = build_oracle_step(workload, batch, threads, beta=beta)
    t_start = time.perf_counter()
    for _ in range(warmup):
        step()
        if budget_s and time.perf_counter() - t_start > budget_s / 3:
            break
    ts = []
    for _ in range(steps_n):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
        if budget_s and len(ts) >= 3 and time.perf_counter() - t_start > budget_s:
            break
    n = len(ts)
    ts.sort()
    med = ts[len(ts) // 2]
    return batch / med, med, n


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the SAME per-step batch as the CUDA arm (args.batch); the sample is bounded in STEPS (wall-clock budget), never
    # by shrinking the batch: a smaller batch is a different configuration (CPU img/s grows with the batch)
    batch = args.batch
    t0 = time.perf_counter()
    ips, med, timed = time_cpu(args.workload, batch, threads, args.warmup, args.steps, beta=args.beta, budget_s=240.0)
    line = {
        "impl": "reference", "metric": "train_img_per_s", "value": round(ips, 3), "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(med * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, batch),
        "cpu_baseline": {"value": round(ips, 3), "unit": "img/s", "cores": threads, "kind": "port",
                         "sample": f"{timed} timed steps (median; {args.steps} requested, 240 s wall budget) of the "
                                   f"same step at batch {batch} on the host cores, oracle/steps.py restating the "
                                   f"reference loop (experiments/new_betavaegan.py:93-193 / new_gan.py:84-128 / "
                                   f"new_vae.py:53-60); wall {time.perf_counter() - t0:.0f}s"},
        "e2e": {"value": round(ips, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def profile_traffic(args):
    """DRAM bytes (read + write) per launch of the GEMM-class kernel from the committed ncu capture of this workload
    (profiles/r02_traffic.json, written by tools/summarize_ncu_metrics.py from the ncu launch list of one step:
    dram__bytes_read.sum + dram__bytes_write.sum over the GEMM-class launches); None for configurations not captured."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(f"{args.workload}_b{args.batch}")
    return None


def workload_config(args, per_gpu_batch, gpus=None):
    gpus = args.gpus if gpus is None else gpus
    names = {"betavaegan": f"VAE-GAN baseline (Larsen, Dis_l loss; experiments/new_betavaegan.py step, beta={args.beta:g})",
             "gan": "GAN (experiments/new_gan.py step)", "vae": "VAE (experiments/new_vae.py step)"}
    return {"workload": names[args.workload] + f", 64x64x3, batch {per_gpu_batch}/GPU",
            "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * gpus, "beta": args.beta,
            "parallelism": f"dp{gpus}",
            "l2": "no explicit flush: every step streams ~2 GB of parameters + Adam state, far beyond the 126 MB L2",
            "gflop_per_img_algorithmic": GFLOP_PER_IMG[args.workload]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from disentangle_mlp_b200 import _lib, ops
    from disentangle_mlp_b200 import model as dm
    from disentangle_mlp_b200 import trainer as tr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    b = args.batch
    torch.manual_seed(999)
    np.random.seed(999)  # identical label stream on every rank
    opt = dm.default_opt()
    if args.workload == "betavaegan":
        eg, d = dm.VAE(opt), dm.Discriminator_celeba(opt)
        eg.apply(dm.weights_init)
        d.apply(dm.weights_init)
        T = tr.BetaVAEGANTrainer(eg.to(dev), d.to(dev), beta=args.beta, lr=1e-3)
        key = "recon_enc"
    elif args.workload == "gan":
        g, d = dm.Generator_celeba(opt), dm.Discriminator_celeba(opt)
        g.apply(dm.weights_init)
        d.apply(dm.weights_init)
        T = tr.GANTrainer(g.to(dev), d.to(dev), lr=3e-4)
        key = "errG"
    else:
        m = dm.VAE(opt)
        m.apply(dm.weights_init)
        T = tr.VAETrainer(m.to(dev), lr=3e-4)
        key = "loss"

    if not args.no_graph:
        # whole-step CUDA graph: one launch per step instead of ~480 (for world > 1 the NCCL all-reduces are captured
        # too; if that fails on this stack the trainer keeps launching eagerly)
        try:
            T.enable_graph(b)  # fp32 NCHW batches already resident in HBM: the `value` measurement
            T.enable_graph(b, input_u8=True)  # uint8 NHWC batches from pinned host memory: the `e2e` measurement
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            T._graph = None
            if rank == 0:
                print(f"[bench] graph capture with NCCL failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)

    torch.manual_seed(999 + 7919 * (rank + 1))  # identical initial weights above, per-rank noise / eps streams below
    # synthetic CelebA-shaped inputs, U[-1,1]; a pool of distinct batches, per-rank seed
    npool = 8
    gen = torch.Generator().manual_seed(1234 + rank)
    dev_pool = [(torch.rand(b, 3, 64, 64, generator=gen) * 2 - 1).to(dev) for _ in range(npool)]
    # end-to-end input format = what a loader over pre-decoded CelebA shards yields: uint8 NHWC in pinned host memory
    # (SURVEY.md 8-f2); ToTensor + Normalize(.5,.5) (dataloader/dataset.py:37-43) runs inside the step's first kernel
    host_pool = [torch.randint(0, 256, (b, 64, 64, 3), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(npool)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsteps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def step_resident(i):
        T.step(dev_pool[i % npool])

    sink = []
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    pending = []

    def drain(keep):
        while len(pending) > keep:
            j = pending.pop(0)
            loss_ev[j % 2].synchronize()
            sink.append(float(loss_host[j % 2]))  # the host now holds step j's loss

    def step_e2e(i):
        # H2D from pinned memory inside the timed region (graph mode copies straight into the static input buffer);
        # every step's loss is copied device -> pinned host and read by the host, one step behind the GPU (the read of
        # step i happens while step i+1 runs), as a training loop that logs its losses does
        x = host_pool[i % npool] if T._graph is not None else host_pool[i % npool].to(dev, non_blocking=True)
        m = T.step(x)
        drain(1)
        loss_host[i % 2].copy_(m[key], non_blocking=True)
        loss_ev[i % 2].record()
        pending.append(i)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # (returns once nvidia-smi has printed its first sample)
    for i in range(max(3, args.warmup)):
        step_resident(i)
    torch.cuda.synchronize()
    if rank == 0:
        sampler.rows.clear()  # keep only samples taken from here on: the timed regions
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    graph_mode = T._graph is not None
    if graph_mode:  # replayed kernels do not pass through the C ABI again: count what the graph holds
        launches = T.graph_launches_per_step * args.steps
    for i in range(2):
        step_e2e(i)
    drain(0)

    def e2e_run(n):
        for i in range(n):
            step_e2e(i)
        drain(0)  # the last loss is read inside the timed region too

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    assert len(sink) == args.steps + 2 and all(v == v for v in sink), "e2e: a loss was not read back (or is NaN)"
    clocks = sampler.stop() if rank == 0 else None  # sampled across BOTH timed regions (device-resident and e2e)

    # roofline pass: same workload launched eagerly (events cannot be read back from a replayed graph), per-launch
    # CUDA events around every GEMM-class kernel on the launching stream
    saved_graph, T._graph = T._graph, None
    T.BIG_ADAM = "now"  # (no side-stream Adam under the GEMMs while they are being timed one by one)
    from disentangle_mlp_b200 import engine as _engine

    saved_wgrad_stream, _engine.WgradSide.stream = _engine.WgradSide.stream, None  # (nor side-stream weight gradients)
    ops.profile_enable(True)
    ops.profile_read()
    t_prof = timed(step_resident, args.steps)
    # per-launch records (kind: 0 dense GEMM = Linear layers and the 3-channel col GEMMs, 1-3 = 5x5 conv forward /
    # transposed / weight gradient) -> totals for the whole GEMM class and for the convolutions alone
    import csv
    import tempfile

    with tempfile.NamedTemporaryFile("r", suffix=".csv") as tf:
        _lib.check(_lib.load().dm_profile_dump(tf.name.encode()), "dm_profile_dump")
        recs = list(csv.DictReader(open(tf.name)))
    gemm_ms = sum(float(r["us"]) for r in recs) / 1e3
    gemm_flops = sum(float(r["gflop"]) for r in recs) * 1e9
    gemm_n = len(recs)
    conv = [r for r in recs if r["kind"] in ("1", "2", "3")]  # the >= 32-channel 5x5 layers (tensor-bound)
    conv3 = [r for r in recs if r["kind"] == "4"]  # the three 3-image-channel layers (TMA-row / HBM bound)
    conv_ms = sum(float(r["us"]) for r in conv) / 1e3
    conv_flops = sum(float(r["gflop"]) for r in conv) * 1e9
    # the same pass with the BatchNorm statistics in their own kernels instead of the GEMM epilogues: isolates the
    # tensor-core part of the GEMM-class kernel (the timed runs above use the fused form: it is the faster STEP)
    nofuse = None
    if _engine.FUSE_BN_STATS:
        _engine.FUSE_BN_STATS = False
        timed(step_resident, 2)
        ops.profile_read()
        timed(step_resident, args.steps)
        with tempfile.NamedTemporaryFile("r", suffix=".csv") as tf:
            _lib.check(_lib.load().dm_profile_dump(tf.name.encode()), "dm_profile_dump")
            recs2 = list(csv.DictReader(open(tf.name)))
        _engine.FUSE_BN_STATS = True
        conv2 = [r for r in recs2 if r["kind"] in ("1", "2", "3")]
        nofuse = {"gemm_ms": sum(float(r["us"]) for r in recs2) / 1e3, "gemm_flops": sum(float(r["gflop"]) for r in recs2) * 1e9,
                  "conv_ms": sum(float(r["us"]) for r in conv2) / 1e3, "conv_flops": sum(float(r["gflop"]) for r in conv2) * 1e9}
    ops.profile_enable(False)
    _engine.WgradSide.stream = saved_wgrad_stream
    T._graph = saved_graph

    if rank == 0:
        peaks = load_peaks()
        gbatch = b * world
        ips = gbatch * args.steps / (ms / 1e3)
        ips_e2e = gbatch * args.steps / (ms_e2e / 1e3)
        ach = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        # denominator: the BURST figure -- the timed region is ~0.1 s at full clocks with no power cap;
        # the fraction of the sustained figure is printed next to it
        peak = peaks["bf16_burst"]
        peak_sus = peaks["bf16_sustained"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cb = b
            t0 = time.perf_counter()
            cips, cmed, _ = time_cpu(args.workload, cb, threads, 1, 3, beta=args.beta)
            cpu = {"value": round(cips, 3), "unit": "img/s", "cores": threads, "kind": "port",
                   "sample": f"3 timed steps (median {cmed:.2f}s) + 1 warm-up of the same step at batch {cb}, "
                             f"oracle/steps.py on the host cores, {time.perf_counter() - t0:.0f}s wall"}
        torch_gpu = None
        if world == 1 and not args.no_torch_baseline:
            torch_gpu = time_torch_gpu(args.workload, b, args.beta)
        line = {
            "metric": "train_img_per_s", "value": round(ips, 2), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, b),
            "clocks": clocks,
            "e2e": {"value": round(ips_e2e, 2), "unit": "img/s", "h2d_bytes_per_step": b * 3 * 64 * 64,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 4),
                    "input": "uint8 NHWC [b,64,64,3] batches in pinned host memory -> H2D -> normalisation fused into "
                             "the step's first kernel; the step's loss is read back every step"},
            "gpu_launches": int(launches), "cuda_graph": bool(graph_mode),
            "roofline": {"bound": "tensor", "kernel": "dm_tapgemm_kernel (all GEMM-class launches of the step)",
                         "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(ach / peak, 4) if peak else None, "traffic": profile_traffic(args),
                         "frac_of_sustained": round(ach / peak_sus, 4) if peak_sus else None,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops = burst ({peaks['src']}); sustained {peak_sus}",
                         "launches_per_step": gemm_n / args.steps, "gemm_ms_per_step": round(gemm_ms / args.steps, 4),
                         # both from the SAME eagerly launched pass (the graph replay has no per-kernel events);
                         # the CUPTI timeline of the replay and the ncu launch list give 0.41-0.44 (profiles/)
                         "gemm_share_of_step": round(gemm_ms / t_prof, 4),
                         "timing": "CUDA events around every GEMM-class launch on the launching stream",
                         "conv_gemms": {"achieved": round(conv_flops / (conv_ms / 1e3) / 1e12, 2) if conv_ms > 0 else None,
                                        "frac": round(conv_flops / (conv_ms / 1e3) / 1e12 / peak, 4) if conv_ms > 0 else None,
                                        "launches_per_step": len(conv) / args.steps,
                                        "ms_per_step": round(conv_ms / args.steps, 4),
                                        "note": "the 5x5 conv / transposed-conv / weight-gradient launches of the layers "
                                                "with >= 32 channels on both sides (north-star: fraction of dense-bf16 "
                                                "peak on the conv GEMMs); the 3-image-channel layers are listed apart"},
                         "conv3_gemms": {"launches_per_step": len(conv3) / args.steps,
                                         "ms_per_step": round(sum(float(r["us"]) for r in conv3) / 1e3 / args.steps, 4),
                                         "gflop_per_step": round(sum(float(r["gflop"]) for r in conv3) / args.steps, 2),
                                         "note": "convs.0 / features.0 / deconv4 gradients: window GEMMs over the padded "
                                                 "3-channel image, 0.4 % of the step's FLOPs, bound by TMA rows and HBM"},
                         "bn_stats_in_epilogue": bool(nofuse is not None),
                         "without_fused_bn_stats": None if nofuse is None else {
                             "note": "same eager pass with the BatchNorm statistics + finalize in separate kernels "
                                     "(DM_BN_FUSE_GEMM=0): the GEMM-class kernel's tensor-core part alone; the fused form "
                                     "above carries that HBM-bound work inside the GEMM launches and is the faster step",
                             "achieved": round(nofuse["gemm_flops"] / (nofuse["gemm_ms"] / 1e3) / 1e12, 2),
                             "frac": round(nofuse["gemm_flops"] / (nofuse["gemm_ms"] / 1e3) / 1e12 / peak, 4),
                             "gemm_ms_per_step": round(nofuse["gemm_ms"] / args.steps, 4),
                             "conv_gemms": {"achieved": round(nofuse["conv_flops"] / (nofuse["conv_ms"] / 1e3) / 1e12, 2),
                                            "frac": round(nofuse["conv_flops"] / (nofuse["conv_ms"] / 1e3) / 1e12 / peak, 4),
                                            "ms_per_step": round(nofuse["conv_ms"] / args.steps, 4)}},
                         "step_frac_of_peak": round(ips / world * GFLOP_PER_IMG[args.workload] / 1e3 / peak, 4)},
            "cpu_baseline": cpu,
            "torch_gpu_baseline": torch_gpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing down NCCL communicators that a captured CUDA graph still references (observed to hang
        # at interpreter exit): everything has been printed and synchronised, so exit hard on every rank
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="betavaegan", choices=["betavaegan", "gan", "vae"])
    ap.add_argument("--beta", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock-torch-on-this-GPU baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

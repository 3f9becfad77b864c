"""Fused training steps: one iteration of the reference's three loops, scheduled as kernel calls.

  VAETrainer.step          <- experiments/new_vae.py:53-60
  GANTrainer.step          <- experiments/new_gan.py:84-128
  BetaVAEGANTrainer.step   <- experiments/new_betavaegan.py:93-193

Semantics kept from the reference: same update order (D, then EG "decoder" phase, then
EG "encoder" phase), both EG Adam steps update every encoder and decoder parameter, `fake` is generated
before the D update and re-scored after it, one (real, fake) label pair per step, BatchNorm always in
training mode with running stats updated on every forward (D 5x, encoder 2x, decoder 3x per step), BCE is a
mean over the batch, MSE / Dis_l / KL are sums.
What is NOT repeated: the reference walks the same graphs with six separate `.backward()` calls and computes
discriminator weight gradients in the EG phases that the next `netD.zero_grad()` throws away
(new_betavaegan.py:95,157-163); here each phase runs one backward per graph and skips the discarded wgrads.
The resulting parameter updates are the same sums of the same terms.

Data parallel (one process per GPU): gradients are SUM-allreduced before each Adam step; the BCE terms are
scaled by 1/world_size so that the sum reproduces the reference's mean over the global batch; BatchNorm
statistics stay per rank (as under the reference's nn.DataParallel).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine, ops
from .ops import BF16, F32

ALIGN = 64  # elements; keeps every parameter's bf16 shadow 128-byte aligned for TMA
BIG_LINEAR = ("lth_features.0.weight", "x_to_mu.0.weight", "x_to_logvar.0.weight", "preprocess.0.weight")
EARLY_BUCKETS = ("lth_features.0.weight", "x_to_mu.0.weight", "x_to_logvar.0.weight")  # all-reduced as soon as final


def _complement(ranges, total):
    """Sorted gaps of [0, total) not covered by the sorted, disjoint `ranges`."""
    out, lo = [], 0
    for a, b in sorted(ranges):
        if a > lo:
            out.append((lo, a))
        lo = b
    if lo < total:
        out.append((lo, total))
    return out


def plan_layout(named_shapes, bf16_big=True):
    """Placement of a module's parameters in the flat buffers and the ranges the optimizer / all-reduce work on.
    Pure host logic (no tensors): `named_shapes` = [(name, shape)] in named_parameters() order.

    Everything is placed in parameter order EXCEPT the three 33.5 M-element Linear weights (EARLY_BUCKETS), which go
    last: the small parameters then form ONE contiguous range [0, small_end) -- one phase-end all-reduce instead of
    three -- with the decoder's parameters as its tail (FlatParams.reduce_from)."""
    numel = {n: int(np.prod(sh)) if len(sh) else 1 for n, sh in named_shapes}
    al = lambda k: (k + ALIGN - 1) // ALIGN * ALIGN  # noqa: E731
    names = [n for n, _ in named_shapes]
    order = [n for n in names if n not in EARLY_BUCKETS] + [n for n in names if n in EARLY_BUCKETS]
    offsets, off = {}, 0
    for n in order:
        offsets[n] = off
        off += al(numel[n])
    plan = {"names": names, "offsets": offsets, "total": off, "numel": numel,
            "small_end": sum(al(numel[n]) for n in names if n not in EARLY_BUCKETS)}
    # 5x5 conv / deconv weights [cs][cb][5][5] with 32-aligned channel counts: tap-major storage (see FlatParams)
    plan["packed"] = {n for n, sh in named_shapes if len(sh) == 4 and sh[0] % 32 == 0 and sh[1] % 32 == 0}
    # gradients kept, all-reduced and read by Adam in bf16
    plan["big16"] = [n for n in names if n in EARLY_BUCKETS] if bf16_big else []
    plan["off16"], o16 = {}, 0
    for n in plan["big16"]:
        plan["off16"][n] = o16
        o16 += numel[n]
    plan["total16"] = o16
    # Adam runs per segment: fp32-gradient stretches of the flat buffer, and the bf16-gradient tensors
    segs, lo = [], 0
    for n in sorted(plan["big16"], key=lambda q: offsets[q]):
        a, k = offsets[n], numel[n]
        if a > lo:
            segs.append((lo, a, None))
        segs.append((a, a + k, plan["off16"][n]))
        lo = a + k
    if lo < off:
        segs.append((lo, off, None))
    plan["segments"] = segs
    # zero_grad() skips the big Linear gradients (written in overwrite mode by the first backward of a phase)
    plan["zero_ranges"] = _complement([(offsets[n], offsets[n] + numel[n]) for n in names if n in BIG_LINEAR], off)
    # data parallel: the three 33.5 M-element gradients are all-reduced early, the remainder at phase end
    plan["late_ranges"] = _complement([(offsets[n], offsets[n] + numel[n]) for n in names if n in EARLY_BUCKETS], off)
    return plan


class FlatParams:
    """Re-homes a module's parameters into ONE flat fp32 buffer (parameters become views), with matching
    flat gradient / Adam-moment buffers and a bf16 shadow that the GEMMs read."""

    def __init__(self, module: torch.nn.Module, lr, betas=(0.9, 0.999), eps=1e-8):
        named = list(module.named_parameters())
        self.names = [n for n, _ in named]
        dev = named[0][1].device
        assert dev.type == "cuda", "FlatParams needs the module on a CUDA device"
        plan = plan_layout([(n, tuple(p.shape)) for n, p in named], os.environ.get("DM_BF16_BIGGRAD", "1") != "0")
        self.offsets, self.total, self._small_end = plan["offsets"], plan["total"], plan["small_end"]
        off = self.total
        self.flat = torch.zeros(off, dtype=F32, device=dev)
        self.grad = torch.zeros(off, dtype=F32, device=dev)
        self.m = torch.zeros(off, dtype=F32, device=dev)
        self.v = torch.zeros(off, dtype=F32, device=dev)
        self.shadow = torch.zeros(off, dtype=BF16, device=dev)
        self.P, self.G, self.W16, self.GP = {}, {}, {}, {}
        self._tail_done = None
        # 5x5 conv / deconv weights [cs][cb][5][5] with 32-aligned channel counts live in the flat buffers in the
        # TAP-MAJOR layout [5][5][cs][cb]: that is the layout of the GEMM operand (the bf16 shadow IS w_down) and of
        # the packed weight gradient the wgrad kernel reduces into (bulk tensor reductions need unit inner stride).
        # The nn.Parameter becomes a permuted (non-contiguous) view with the reference shape; Adam is elementwise,
        # so it does not care.
        self.packed = plan["packed"]
        for n, p in named:
            o, k = self.offsets[n], p.numel()
            self._layout(self.flat, n, p.shape).copy_(p.data)
            p.data = self._layout(self.flat, n, p.shape)
            self.P[n] = p
            self.G[n] = self._layout(self.grad, n, p.shape)
            self.W16[n] = self._layout(self.shadow, n, p.shape)
            if n in self.packed:
                self.GP[n] = self.grad[o:o + k].view(25, p.shape[0], p.shape[1])
        # The three 16384x2048 Linear weight gradients (92 % of all gradient elements) are produced ONCE per phase by
        # a GEMM whose epilogue can store bf16 directly: they are kept, all-reduced and read by Adam in bf16 (half the
        # NCCL bytes, 2 bytes less per parameter in the weight-gradient store and in the Adam read).  The fp32
        # accumulation happens in TMEM; one rounding to bf16 (2^-9) is far below the bf16 activation noise.
        self.big16, self.off16 = plan["big16"], plan["off16"]
        self.grad16 = torch.zeros(max(plan["total16"], 1), dtype=BF16, device=dev)
        for n in self.big16:
            self.G[n] = self.grad16[self.off16[n]:self.off16[n] + self.P[n].numel()].view(self.P[n].shape)
        self._segments = plan["segments"]  # Adam runs per segment (fp32-gradient stretches / bf16-gradient tensors)
        self.module = module
        self.buffers = dict(module.named_buffers())
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.step_dev = torch.zeros((), dtype=torch.int32, device=dev)  # the kernels' copy of step_count
        self.cache = engine.OperandCache()
        self.cache.lin_views = self.W16
        # persistent bf16 operand packs of the conv / deconv weights ([cs][cb][5][5]), refreshed IN PLACE after every
        # update: no lazily rebuilt state, so a captured CUDA graph always reads current operands
        self.cache.static_packs = {}
        self._up_std = {}
        self.cache.packed_grads = self.GP
        for n, p in named:
            if p.dim() == 4:
                cs, cb = p.shape[0], p.shape[1]
                if n in self.packed:
                    o, k = self.offsets[n], p.numel()
                    w_down = self.shadow[o:o + k].view(25, cs, cb)
                    w_up = torch.empty((25, cb, cs), dtype=BF16, device=dev)
                    if ops.up_merged(cb, 2):  # conv_up reads the phase-merged pack [9][4cb][cs]; w_up is its source
                        self._up_std[n[:-len(".weight")]] = w_up
                        w_up = torch.empty((9, 4 * cb, cs), dtype=BF16, device=dev)
                    w_pair = torch.empty((15, cs, 64), dtype=BF16, device=dev) if ops.down_paired(cb, 2) else None
                    self.cache.static_packs[n[:-len(".weight")]] = (w_down, w_up, w_pair)
                else:
                    self.cache.static_packs[n[:-len(".weight")]] = engine.pack3(
                        p.detach(), engine.CONV3_STRIDE[n[:-len(".weight")]])
        self._zero_ranges, self._late_ranges = plan["zero_ranges"], plan["late_ranges"]
        self.reducer, self.shard, self._gather_pending = None, False, False
        # deferred update of the big bf16-gradient tensors (adam(big="defer")): device flag "grad16 holds an unapplied
        # gradient" (read by the gated Adam launch inside the step graph) + its host mirror
        self.big_valid = torch.zeros((), dtype=torch.int32, device=dev)
        self._big_pending, self._big_event = False, None
        self.params_changed()
        module.register_load_state_dict_pre_hook(lambda *_a, **_k: self.flush())
        module.register_state_dict_pre_hook(lambda *_a, **_k: self.flush())  # (direct .parameters() reads: T.sync() first)
        # module.load_state_dict() copies into the re-homed fp32 masters: refresh the bf16 shadow / operand packs
        module.register_load_state_dict_post_hook(lambda _m, _keys: self.params_changed())

    def _layout(self, buf, n, shape):
        """View of parameter `n`'s slice of a flat buffer with the parameter's (reference) shape."""
        o = self.offsets[n]
        k = int(np.prod(shape)) if len(shape) else 1
        if n in getattr(self, "packed", ()):
            cs, cb = shape[0], shape[1]
            return buf[o:o + k].view(5, 5, cs, cb).permute(2, 3, 0, 1)
        return buf[o:o + k].view(shape)

    def attach(self, reducer):
        """Data parallel: the three 16384x2048 Linear weights (92 % of all parameters) get a SHARDED optimizer step
        (ZeRO-1 on those tensors): their bf16 gradients are reduce-scattered, every rank runs Adam on 1/world of each
        tensor (fp32 master, m, v of its chunk) and the bf16 shadow -- what the GEMMs read -- is all-gathered.  Same wire
        bytes as the all-reduce it replaces, but Adam's HBM traffic drops by (world-1)/world on 92 % of the parameters.
        The fp32 masters / moments of the other ranks' chunks go stale: gather_masters() before reading them
        (checkpoints).  DM_SHARD_BIG=0: replicated Adam + all-reduce."""
        self.reducer = reducer
        self.shard = bool(reducer.on and self.big16 and os.environ.get("DM_SHARD_BIG", "1") != "0"
                          and all(self.P[n].numel() % reducer.world == 0 for n in self.big16))

    def reduce_early(self, reducer, name):
        if name in self.off16:
            o = self.off16[name]
            if self.shard:
                reducer.reduce_scatter_async(self.grad16, o, self.P[name].numel())
                self._rs_event = reducer.mark() if self.grad16.is_cuda else None
            else:
                reducer.allreduce_async(self.grad16, o, o + self.P[name].numel())
            return
        o = self.offsets[name]
        reducer.allreduce_async(self.grad, o, o + self.P[name].numel())

    def reduce_from(self, reducer, name):
        """Data parallel: all-reduce the gradients of parameter `name` and everything after it in the flat buffer NOW
        (they are final), on the side stream; reduce_rest() then skips that tail.  Used for the decoder parameters
        (the tail of the VAE's parameter list), whose gradients are final before the encoder backward starts."""
        if not reducer.on:
            return
        lo = self.offsets[name]
        assert lo < self._small_end
        reducer.allreduce_async(self.grad, lo, self._small_end)
        self._tail_done = lo

    def reduce_rest(self, reducer):
        """Enqueue the all-reduce of everything not reduced early; the caller joins with reducer.wait()."""
        stop = self._tail_done if self._tail_done is not None else self.total
        for lo, hi in self._late_ranges:
            if lo < stop:
                reducer.allreduce_async(self.grad, lo, min(hi, stop))
            elif self._tail_done is not None and lo >= self._small_end:
                reducer.allreduce_async(self.grad, lo, hi)  # (a late range behind the early tail: none today)
        self._tail_done = None

    def reduce_rest_and_wait(self, reducer):
        """All-reduce what is still unreduced and order the stream behind the collectives the following adam() needs.
        Sharded: only behind the reduce-scatters of the big gradients -- adam() updates those chunks first and waits for
        the small gradients' all-reduce (which then ran under that HBM-bound work) right before it needs them."""
        self.reduce_rest(reducer)
        ev = getattr(self, "_rs_event", None)
        if self.shard and ev is not None and os.environ.get("DM_SPLIT_WAIT", "1") != "0":
            torch.cuda.current_stream().wait_event(ev)
            self._rs_event = None
            self._wait_rest = reducer.mark()  # (an event behind the all-reduce just enqueued, not behind later gathers)
        else:
            reducer.wait()

    def refresh_packs(self):
        for name, packs in self.cache.static_packs.items():
            p = self.P[name + ".weight"]
            if name + ".weight" in self.packed:  # w_down is the shadow itself; w_up = per-tap transpose of it
                if name in self._up_std:
                    ops.transpose(packs[0], 25, p.shape[0], p.shape[1], out=self._up_std[name])
                    ops.pack_up_merged(self._up_std[name], p.shape[0], p.shape[1], out=packs[1])
                else:
                    ops.transpose(packs[0], 25, p.shape[0], p.shape[1], out=packs[1])
                if packs[2] is not None:
                    ops.pack_down_pairs(packs[0], p.shape[0], p.shape[1], out=packs[2])
            else:
                engine.pack3(p.detach(), engine.CONV3_STRIDE[name], out=packs)

    def touch(self):
        """The fp32 masters were updated through raw pointers (Adam kernel, graph replay, restore): nn.Parameter
        version counters did not move, so the MODULE's own lazily built operand cache (engine.OperandCache, used by
        the drop-in forward path: decode / model(x) sampling between training steps) must be dropped explicitly."""
        c = getattr(self.module, "_operand_cache", None)
        if c is not None:
            c.invalidate()

    def params_changed(self):
        """Call after the fp32 parameters were modified outside adam() (init, load_state_dict)."""
        ops.cast_bf16(self.flat, self.shadow)
        self.refresh_packs()
        self.touch()

    def zero_grad(self):
        for lo, hi in self._zero_ranges:
            self.grad[lo:hi].zero_()

    def adam(self, grad_scale=1.0, gather=True, big="now", side=None):
        """One Adam update; the step count lives on the device (incremented by the kernel) so that the call can be
        replayed from a CUDA graph; the host mirror `step_count` is kept for the optimizer state dict.
        Sharded big tensors (attach()): only this rank's chunk is updated; gather=True launches the all-gather of the
        bf16 shadows right away (asynchronously on the NCCL stream: reducer.wait() before their first use),
        gather=False leaves it pending for gather_if_pending() -- e.g. at the start of the next step, under work that
        does not read those weights.
        big (not sharded): what to do with the big bf16-gradient tensors (2 x 33.5 M elements = 2/3 of this call's HBM
        traffic for the VAE): "now" = inline; "side" = on the stream `side` right away, wait_big() before their first
        use; "defer" = leave the gradient in place, flagged valid, and apply it in finish_big() -- at the start of the
        next step on a side stream, under work that does not read those weights (flush() applies it immediately)."""
        self.step_count += 1
        early = getattr(self, "_big_early", False)  # adam_big_early() has already updated them in this step
        self._big_early = False
        if early:  # join the side stream (the update had the rest of the backward pass to finish)
            self.wait_big()
            self._big_event = None
        split = (big != "now" or early) and not self.shard and bool(self.big16)
        # sharded: the big tensors' chunks first, each followed at once by the all-gather of its bf16 shadow -- the NCCL
        # stream then works under the rest of this call (small tensors, operand packs) and the next phase's convolutions
        order = sorted(range(len(self._segments)), key=lambda i: (self._segments[i][2] is None) if self.shard else i)
        eager_gather = self.shard and gather and self.shadow.is_cuda and os.environ.get("DM_EAGER_GATHER", "1") != "0"
        wait_rest, self._wait_rest = getattr(self, "_wait_rest", None), None
        for k, i in enumerate(order):
            lo, hi, o16 = self._segments[i]
            if o16 is not None and split:
                continue
            if o16 is None and wait_rest is not None:  # (reduce_rest_and_wait: the small gradients are needed from here)
                torch.cuda.current_stream().wait_event(wait_rest)
                wait_rest = None
            full = (lo, hi)
            if o16 is not None and self.shard:
                a, b = self.reducer.chunk(hi - lo)
                g = self.grad16[o16 + a:o16 + b]
                lo, hi = lo + a, lo + b
            else:
                g = self.grad[lo:hi] if o16 is None else self.grad16[o16:o16 + (hi - lo)]
            ops.adam_step(self.flat[lo:hi], g, self.m[lo:hi], self.v[lo:hi], self.lr, self.betas[0], self.betas[1],
                          self.eps, 0, grad_scale, self.shadow[lo:hi], step_dev=self.step_dev, count_step=(k == 0))
            if o16 is not None and eager_gather:
                self.reducer.all_gather_async(self.shadow, full[0], full[1] - full[0])
        if wait_rest is not None:
            torch.cuda.current_stream().wait_event(wait_rest)
        if eager_gather:
            self._gather_pending = False
            self._gather_event = self.reducer.mark()
        self.refresh_packs()  # bf16 conv operand packs follow the updated fp32 weights
        self.touch()
        if self.shard and not eager_gather:
            self._gather_pending = True
            if gather:
                self.gather_if_pending()
        if split and not early:
            self.big_valid.fill_(1)
            self._big_pending = True
            if big == "side":
                self.finish_big(side)

    def adam_big_early(self, side):
        """The update of the big bf16-gradient tensors, launched as soon as their gradients are final (the encoder's
        heads have been back-propagated) on stream `side`, under the rest of the backward pass; the adam() call that
        ends the phase then covers everything else.  Uses step counter + 1: adam() has not incremented it yet."""
        if not self.big16 or self.shard:
            return
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for lo, hi, o16 in self._segments:
                if o16 is not None:
                    ops.adam_step(self.flat[lo:hi], self.grad16[o16:o16 + (hi - lo)], self.m[lo:hi], self.v[lo:hi], self.lr,
                                  self.betas[0], self.betas[1], self.eps, 0, 1.0, self.shadow[lo:hi],
                                  step_dev=self.step_dev, step_offset=1)
            self._big_event = torch.cuda.Event()
            self._big_event.record(side)
        self._big_early = True

    def finish_big(self, side=None):
        """Apply the pending update of the big tensors (gated by the device flag, so the launches are safe to capture
        and replay): on stream `side` when given (asynchronous; wait_big() orders a consumer behind it), else inline."""
        if not self.big16 or self.shard:
            return

        def run():
            for lo, hi, o16 in self._segments:
                if o16 is not None:
                    ops.adam_step(self.flat[lo:hi], self.grad16[o16:o16 + (hi - lo)], self.m[lo:hi], self.v[lo:hi], self.lr,
                                  self.betas[0], self.betas[1], self.eps, 0, 1.0, self.shadow[lo:hi],
                                  step_dev=self.step_dev, enable=self.big_valid)

        if side is None:
            run()
            self._big_event = None
        else:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()
                self._big_event = torch.cuda.Event()
                self._big_event.record(side)
        self._big_pending = False
        self.touch()

    def wait_big(self):
        """Order the current stream behind the last side-stream update of the big tensors."""
        if self._big_event is not None:
            torch.cuda.current_stream().wait_event(self._big_event)

    def flush(self):
        """Apply any deferred update NOW (before the parameters / moments are read or replaced from outside)."""
        if self._big_pending or self._big_event is not None:
            self.wait_big()
            if self._big_pending:
                self.finish_big()
            self.big_valid.zero_()  # a captured step graph must not apply this gradient again
            self._big_event = None

    def gather_if_pending(self):
        """All-gather the bf16 shadows of the sharded tensors (asynchronous, NCCL stream) if an update is pending."""
        if not (self.shard and self._gather_pending):
            return
        for n in self.big16:
            self.reducer.all_gather_async(self.shadow, self.offsets[n], self.P[n].numel())
        self._gather_pending = False
        self._gather_event = self.reducer.mark() if self.shadow.is_cuda else None

    def wait_gathered(self):
        """Order the current stream behind the last all-gather of this optimizer's bf16 shadows (and only that)."""
        if not self.shard:
            return
        ev = getattr(self, "_gather_event", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        elif not self.shadow.is_cuda:
            self.reducer.wait()

    def gather_masters(self):
        """Sharded tensors: make the fp32 masters and Adam moments complete on every rank (checkpoint / comparison)."""
        if not self.shard:
            return
        self.gather_if_pending()
        for n in self.big16:
            for buf in (self.flat, self.m, self.v):
                self.reducer.all_gather_async(buf, self.offsets[n], self.P[n].numel())
        self.reducer.wait()

    def snapshot(self):
        self.flush()
        return {"flat": self.flat.clone(), "m": self.m.clone(), "v": self.v.clone(), "step": self.step_count,
                "buffers": {k: b.clone() for k, b in self.buffers.items()}}

    def restore(self, snap):
        self.flat.copy_(snap["flat"])
        self.m.copy_(snap["m"])
        self.v.copy_(snap["v"])
        self.step_count = snap["step"]
        self.step_dev.fill_(snap["step"])
        for k, b in self.buffers.items():
            b.copy_(snap["buffers"][k])
        self._gather_pending = False
        self._big_pending, self._big_event = False, None
        self.big_valid.zero_()
        self.params_changed()

    def optimizer_state_dict(self):
        """torch.optim.Adam-compatible state (exp_avg / exp_avg_sq / step per parameter)."""
        self.flush()
        state = {}
        for i, n in enumerate(self.names):
            o, k = self.offsets[n], self.P[n].numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self._layout(self.m, n, self.P[n].shape).contiguous().clone(),
                        "exp_avg_sq": self._layout(self.v, n, self.P[n].shape).contiguous().clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        self.flush()
        steps = set()
        for i, n in enumerate(self.names):
            st = sd["state"].get(i)
            if st is None:
                continue
            o, k = self.offsets[n], self.P[n].numel()
            self._layout(self.m, n, self.P[n].shape).copy_(st["exp_avg"].view(self.P[n].shape))
            self._layout(self.v, n, self.P[n].shape).copy_(st["exp_avg_sq"].view(self.P[n].shape))
            steps.add(int(st["step"]))
        if steps:
            assert len(steps) == 1, "per-parameter step counts differ; not representable"
            self.step_count = steps.pop()
            self.step_dev.fill_(self.step_count)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]


class GradReducer:
    """Gradient all-reduce (SUM) over torch.distributed in bucket slices; a no-op for world size 1.

    CUDA tensors: each bucket is enqueued on a side stream, ordered after the work already queued on the
    current stream, so NCCL overlaps whatever the current stream does next; wait() joins the side stream.
    CPU tensors (gloo, used by the CPU tests): asynchronous work handles."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.on else 1
        self.rank = dist.get_rank() if self.on else 0
        self.comm_stream = None
        self.handles = []
        self.cuda_pending = False

    def bce_scale(self):
        """nn.BCELoss is a MEAN over the global batch: each rank scales its local mean by 1/world so that the
        SUM all-reduce of gradients reproduces it; the sum-reduced losses (MSE, Dis_l, KL) need no scaling."""
        return 1.0 / self.world

    @staticmethod
    def buckets(n, bucket_elems=8 << 20):
        lo = 0
        while lo < n:
            hi = min(n, lo + bucket_elems)
            yield lo, hi
            lo = hi

    def allreduce_async(self, flat, lo=0, hi=None):
        if not self.on:
            return
        hi = flat.numel() if hi is None else hi
        view = flat[lo:hi]
        if flat.is_cuda:
            self._on_comm_stream(lambda: self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM))
        else:
            self.handles.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, async_op=True))

    def _on_comm_stream(self, fn):
        """Enqueue fn() on the NCCL side stream, ordered after the work already queued on the current stream."""
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            fn()
        self.cuda_pending = True

    def mark(self):
        """An event on the NCCL stream behind everything enqueued so far: lets a consumer wait for THAT work only
        (current_stream().wait_event), not for collectives enqueued later.  None without CUDA work."""
        if self.comm_stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self.comm_stream)
        return ev

    def chunk(self, n):
        """This rank's slice [lo, hi) of a tensor of n elements split evenly over the ranks."""
        c = n // self.world
        return self.rank * c, (self.rank + 1) * c

    def reduce_scatter_async(self, flat, lo, n):
        """In place: this rank's chunk of flat[lo:lo+n] receives the SUM over ranks (the other chunks become garbage
        nobody reads).  Half the wire bytes of an all-reduce; the other half is the all-gather of what Adam writes."""
        if not self.on:
            return
        assert n % self.world == 0
        view = flat[lo:lo + n]
        a, b = self.chunk(n)
        if flat.is_cuda:
            self._on_comm_stream(lambda: self.dist.reduce_scatter_tensor(view[a:b], view, op=self.dist.ReduceOp.SUM))
        else:  # gloo (CPU tests) has no reduce-scatter: all-reduce, the chunk is then what a reduce-scatter leaves
            self.handles.append(self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, async_op=True))

    def all_gather_async(self, flat, lo, n):
        """In place: every rank's chunk of flat[lo:lo+n] is broadcast to all ranks."""
        if not self.on:
            return
        assert n % self.world == 0
        view = flat[lo:lo + n]
        a, b = self.chunk(n)
        if flat.is_cuda:
            self._on_comm_stream(lambda: self.dist.all_gather_into_tensor(view, view[a:b]))
        else:
            mine = view[a:b].clone()
            self.handles.append(self.dist.all_gather_into_tensor(view, mine, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles.clear()
        if self.cuda_pending:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self.cuda_pending = False


def _scalar(dev):
    return torch.zeros((), dtype=F32, device=dev)


def _scalars(n, dev):
    """n zeroed fp32 device scalars from ONE fill (4-byte aligned 0-dim views, which is all the loss kernels need)"""
    buf = torch.zeros(n, dtype=F32, device=dev)
    return tuple(buf[i] for i in range(n))


class _Base:
    """Common trainer plumbing: label stream, BCE helper, optional whole-step CUDA-graph capture.

    Graph mode (`enable_graph(batch)`; world size 1): the step's ~480 kernel launches are captured once into a
    torch.cuda.CUDAGraph and replayed; the batch is copied into a static device buffer, the per-step labels and
    the Adam step counters are read from device memory, noise / eps come from torch's graph-safe CUDA generator."""

    def __init__(self):
        self.dist = GradReducer()
        self.metrics = {}
        self._graph = None
        if os.environ.get("DM_WGRAD_STREAM", "1") != "0" and engine.WgradSide.stream is None and torch.cuda.is_available():
            engine.WgradSide.stream = torch.cuda.Stream()
        if (os.environ.get("DM_BIG_WGRAD_STREAM", "0") != "0" and engine.WgradSide.big_stream is None
                and torch.cuda.is_available()):
            engine.WgradSide.big_stream = torch.cuda.Stream()

    def flat_params(self):
        raise NotImplementedError

    # where the Adam update of the encoder's two 33.5 M-element Linear weights runs (single GPU / unsharded):
    #   "early" (default) on a side stream as soon as their gradients are final, under the encoder's conv backward;
    #   "defer" the phase-final one at the start of the next step under the discriminator phase (needs sync() before
    #   the parameters are read); "now" inline.  Measured in profiles/ (r02 A/B).
    BIG_ADAM = os.environ.get("DM_BIG_ADAM", "early")
    DEFER = BIG_ADAM == "defer"

    def _heads_done(self, fp):
        if self.BIG_ADAM != "early" or fp.shard:
            return None
        return lambda: fp.adam_big_early(self._side())

    def _deferred(self):
        """FlatParams whose step-final Adam leaves the big tensors to the start of the next step"""
        return []

    def _side(self):
        # ONE side stream per process for the big Linear weights: their weight gradients (engine.WgradSide.big_stream)
        # and the early / deferred Adam updates that consume them are ordered by being on the same stream
        if engine.WgradSide.big_stream is not None:
            return engine.WgradSide.big_stream
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream()
        return self._side_stream

    @staticmethod
    def _before_use(fp):
        """engine hook: order the stream behind whatever still updates fp's big Linear weights (the all-gather of a
        sharded update, or a side-stream / deferred Adam) right before a network first reads them"""
        def wait():
            fp.wait_gathered()
            fp.wait_big()
        return wait

    def _attach(self):
        for fp in self.flat_params():
            fp.attach(self.dist)

    def sync(self, masters=False):
        """Data parallel with sharded Adam: complete every pending all-gather (call before the modules are used
        outside step(): sampling, evaluation); masters=True also gathers the fp32 masters / Adam moments (before
        state_dict() / optimizer_state_dict(): checkpoints)."""
        for fp in self.flat_params():
            fp.flush()
            fp.gather_if_pending()
            if masters:
                fp.gather_masters()
        self.dist.wait()

    @staticmethod
    def draw_labels():
        """One (real, fake) label pair per step from numpy's global RNG (new_betavaegan.py:89-90)."""
        fake = float(np.random.choice(a=[0.1, 0.9], p=[0.95, 0.05]))
        real = float(np.random.choice(a=[0.1, 0.9], p=[0.05, 0.95]))
        return real, fake

    def _bce(self, prob, target, loss, stat=None):
        """Mean BCE over the GLOBAL batch and its gradient w.r.t. prob. target: float or 1-element CUDA tensor."""
        b = prob.numel()
        dprob = torch.empty_like(prob)
        ops.bce_const(prob, target, loss, 1.0, n_total=b * self.dist.world, dprob=dprob, stat=stat)
        return dprob

    def _fork_side(self, fn):
        """Run fn() on a side stream ordered after the work queued so far; returns a token for _join_side().
        (Inside graph capture this becomes a parallel branch of the graph.)  DM_SIDE_ADAM=0: run it inline."""
        if os.environ.get("DM_SIDE_ADAM", "1") == "0":
            fn()
            return None
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self._side_stream.wait_stream(cur)
        with torch.cuda.stream(self._side_stream):
            fn()
        return self._side_stream

    @staticmethod
    def _join_side(token):
        if token is not None:
            torch.cuda.current_stream().wait_stream(token)

    @staticmethod
    def _ingest(data, pim):
        """The step's input batch -> (fp32 NCHW image the losses read, its padded bf16 image in `pim`).  data: fp32
        NCHW in [-1,1] (what the reference's loader yields), or uint8 NHWC [b,64,64,3] straight from a pre-decoded shard:
        the loader's ToTensor + Normalize(.5,.5) (dataloader/dataset.py:37-43) then runs inside the same kernel."""
        if data.dtype == torch.uint8:
            _, x = ops.pad_image3(data, pim, want_nchw=True)
            return x
        ops.pad_image3(data, pim)
        return data

    def _early(self, fp):
        """grad_ready hook: all-reduce a big gradient bucket as soon as the backward pass has finished writing it."""
        if not self.dist.on:
            return None
        return lambda name: fp.reduce_early(self.dist, name)

    # ------------------------------------------------------------------ CUDA graph
    def enable_graph(self, batch, input_u8=False):
        """input_u8: the step will be fed uint8 NHWC [batch,64,64,3] batches (normalisation fused into the first
        kernel) instead of fp32 NCHW [-1,1] ones; the captured graph is specific to one input format."""
        if self.dist.on and os.environ.get("DM_GRAPH_DDP", "1") == "0":
            raise RuntimeError("graph capture with collectives disabled (DM_GRAPH_DDP=0)")
        fps = self.flat_params()
        dev = fps[0].flat.device
        self._gx = (torch.zeros(batch, 64, 64, 3, dtype=torch.uint8, device=dev) if input_u8
                    else torch.zeros(batch, 3, 64, 64, device=dev))
        if not hasattr(self, "_glabels"):  # labels / random inputs are shared by the graphs of both input formats
            self._glabels = torch.zeros(2, device=dev)
            self._glabels_ring = [(torch.zeros(2).pin_memory(), torch.cuda.Event()) for _ in range(16)]
            # random inputs of the step (noise / eps), drawn OUTSIDE the graph in the reference's order
            self._grands = [torch.zeros(batch, 128, device=dev) for _ in range(self.n_rands)]
        snaps = [fp.snapshot() for fp in fps]
        rng = torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up outside capture: lazy inits (func attributes, driver entry points)
            self._glabels.copy_(torch.tensor([0.9, 0.1]))
            if input_u8:
                self._gx.random_(0, 256)
            else:
                self._gx.uniform_(-1, 1)
            for _ in range(2):
                self._step_impl(self._gx, self._glabels[0:1], self._glabels[1:2], *self._grands)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        counts0 = [fp.step_count for fp in fps]
        from . import _lib

        l0 = _lib.launch_count()
        self._graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self._graph):
                self._gmetrics = self._step_impl(self._gx, self._glabels[0:1], self._glabels[1:2], *self._grands)
        finally:
            # events recorded while capturing belong to the graph: they must not be waited on by eager code afterwards
            for fp in fps:
                fp._gather_event = None
                fp._big_event = None
                fp._rs_event = None
                fp._wait_rest = None
            self.dist.cuda_pending = False
        self.graph_launches_per_step = _lib.launch_count() - l0  # libdm_b200 kernels captured in one step
        self._adams_per_step = [fp.step_count - c for fp, c in zip(fps, counts0)]
        for fp, s in zip(fps, snaps):  # undo the warm-up / capture-time bookkeeping: training state as before
            fp.restore(s)
        torch.cuda.set_rng_state(rng, dev)
        torch.cuda.synchronize()
        # one captured graph per input format (fp32 NCHW / uint8 NHWC): _graph_step picks by the batch's dtype
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self._graphs[bool(input_u8)] = (self._graph, self._gx, self._gmetrics)

    def _graph_step(self, data, real_label, fake_label, rands=None):
        """Replay the captured step.  rands: optional list of `n_rands` [batch,128] tensors (noise / eps in the
        reference's draw order) copied into the graph's static buffers instead of drawing them (parity tests)."""
        if real_label is None:
            real_label, fake_label = self.draw_labels()
        # the host may run several steps ahead of the GPU: each step's labels get their own pinned slot, reused only
        # after the copy that read it has completed
        k = self._glabel_slot = (getattr(self, "_glabel_slot", -1) + 1) % len(self._glabels_ring)
        host, ev = self._glabels_ring[k]
        ev.synchronize()
        host[0], host[1] = real_label, fake_label
        self._glabels.copy_(host, non_blocking=True)
        ev.record()
        graph, gx, gmetrics = self._graphs.get(data.dtype == torch.uint8) or (None, None, None)
        if graph is None:
            raise RuntimeError(f"no CUDA graph was captured for {data.dtype} input batches "
                               f"(enable_graph(batch, input_u8={data.dtype == torch.uint8}))")
        gx.copy_(data, non_blocking=True)
        if rands is None:
            for r in self._grands:
                r.normal_()
        else:
            assert len(rands) == len(self._grands) and all(r is not None for r in rands), \
                "graph mode: inject all of the step's random inputs or none"
            for dst, src in zip(self._grands, rands):
                dst.copy_(src, non_blocking=True)
        graph.replay()
        for fp, n in zip(self.flat_params(), self._adams_per_step):
            fp.step_count += n
            fp.touch()
            fp._gather_pending = fp.shard  # the step's last sharded update is gathered at the start of the next replay
            fp._big_pending = fp in self._deferred()  # ... and a deferred big-tensor update is applied there
        self.metrics = gmetrics
        return self.metrics


class VAETrainer(_Base):
    """One step of experiments/new_vae.py:53-60 (loss = MSE-sum + KL-sum, :39-48)."""

    def __init__(self, model, lr=3e-3, beta=1.0):
        super().__init__()
        self.model = model
        self.fp = FlatParams(model, lr)
        self.beta = beta
        self._attach()

    n_rands = 1  # eps

    def flat_params(self):
        return [self.fp]

    def _deferred(self):
        return [self.fp] if (self.DEFER and not self.fp.shard) else []

    def step(self, data, eps=None):
        if self._graph is not None:
            return self._graph_step(data, 0.0, 0.0, None if eps is None else [eps])
        return self._step_impl(data, None, None, eps=eps)

    def _step_impl(self, data, real_label=None, fake_label=None, eps=None):
        fp, dev = self.fp, data.device
        b = data.shape[0]
        loss = _scalar(dev)
        fp.zero_grad()
        fp.gather_if_pending()  # sharded Adam: the previous step's update reaches the other ranks under the encoder convs
        if self.DEFER:  # ... single GPU: the previous step's update of the two big Linear weights runs under them
            fp.finish_big(self._side())
        pim = ops.pim_empty(b, dev)
        data = self._ingest(data, pim)
        mu, logvar, Se = engine.encoder_forward(None, fp.P, fp.buffers, fp.cache, True, pim=pim,
                                                before_heads=self._before_use(fp))
        if eps is None:
            eps = torch.randn_like(mu)
        _, z16 = ops.reparam_forward(mu, logvar, eps)
        recon, Sg = engine.decoder_forward(z16, fp.P, fp.buffers, fp.cache, True)
        drecon = torch.empty_like(recon)
        ops.mse_sum(recon, data, loss, 1.0, drecon, 1.0)
        dmu_kl, dlv_kl = torch.empty_like(mu), torch.empty_like(mu)
        ops.kl(mu, logvar, loss, self.beta, dmu_kl, dlv_kl)
        dz = engine.decoder_backward(Sg, drecon, fp.P, fp.G, fp.cache, True, True, overwrite_big=True)
        fp.reduce_from(self.dist, "preprocess.0.weight")
        _, _, dmu, dlv = ops.reparam_backward(dz, logvar, eps, dmu_kl, dlv_kl)
        engine.encoder_backward(Se, dmu, dlv, fp.P, fp.G, fp.cache, True, overwrite_big=True,
                                grad_ready=self._early(fp), heads_done=self._heads_done(fp))
        fp.reduce_rest_and_wait(self.dist)
        fp.adam(gather=False, big="defer" if self.DEFER else "now")
        self.metrics = {"loss": loss}
        return self.metrics


class GANTrainer(_Base):
    """One step of experiments/new_gan.py:84-128."""

    def __init__(self, netG, netD, lr=3e-3):
        super().__init__()
        self.netG, self.netD = netG, netD
        self.fg = FlatParams(netG, lr)
        self.fd = FlatParams(netD, lr)
        self._attach()

    n_rands = 1  # noise

    def flat_params(self):
        return [self.fg, self.fd]

    def step(self, data, real_label=None, fake_label=None, noise=None):
        if self._graph is not None:
            return self._graph_step(data, real_label, fake_label, None if noise is None else [noise])
        if real_label is None:
            real_label, fake_label = self.draw_labels()
        return self._step_impl(data, real_label, fake_label, noise=noise)

    def _step_impl(self, data, real_label, fake_label, noise=None):
        fg, fd, dev = self.fg, self.fd, data.device
        b = data.shape[0]
        errD, errG, sum_dx, sum_dgz1, sum_dgz2 = _scalars(5, dev)
        # ---- (1) discriminator: real batch, then detached fake batch (:84-113).  Both forward passes go through every
        # GEMM together (stacked along the batch); BatchNorm statistics / running stats stay per pass, real first.
        fd.zero_grad()
        if noise is None:
            noise = torch.randn(b, 128, device=dev)
        # D's inputs as padded bf16 images, stacked [data | fake]: every producer writes its slice (no torch.cat)
        pim = ops.pim_empty(2 * b, dev)
        data = self._ingest(data, pim[:b])
        fake, Sg = engine.decoder_forward(noise, fg.P, fg.buffers, fg.cache, True, pim_out=pim[b:])
        prob, _, S12 = engine.discriminator_forward(None, fd.P, fd.buffers, fd.cache, True, groups=2, pim=pim)
        dprob = torch.empty_like(prob)
        nt = b * self.dist.world
        ops.bce_const(prob[:b], real_label, errD, 1.0, n_total=nt, dprob=dprob[:b], stat=sum_dx)
        ops.bce_const(prob[b:], fake_label, errD, 1.0, n_total=nt, dprob=dprob[b:], stat=sum_dgz1)
        engine.discriminator_backward(S12, dprob, None, fd.P, fd.G, fd.cache, False, True, overwrite_big=True,
                                      grad_ready=self._early(fd))
        del S12
        fd.reduce_rest_and_wait(self.dist)
        fd.adam()
        # ---- (2) generator: re-score the same fake batch with the updated D (:118-128)
        fg.zero_grad()
        prob_g, _, S3 = engine.discriminator_forward(None, fd.P, fd.buffers, fd.cache, True, pim=pim[b:],
                                                     before_linear=fd.wait_gathered if fd.shard else None)
        d3 = self._bce(prob_g, real_label, errG, sum_dgz2)
        dfake = engine.discriminator_backward(S3, d3, None, fd.P, None, fd.cache, True, False)
        engine.decoder_backward(Sg, dfake, fg.P, fg.G, fg.cache, False, True, overwrite_big=True)
        fg.reduce_rest_and_wait(self.dist)
        fg.adam()
        self.metrics = {"errD": errD, "errG": errG, "D_x": sum_dx / b, "D_G_z1": sum_dgz1 / b, "D_G_z2": sum_dgz2 / b}
        return self.metrics


class BetaVAEGANTrainer(_Base):
    """One step of experiments/new_betavaegan.py:93-193 (beta multiplies only the KL term, :64-65)."""

    def __init__(self, netEG, netD, beta, lr=1e-3):
        super().__init__()
        self.netEG, self.netD = netEG, netD
        self.feg = FlatParams(netEG, lr)
        self.fd = FlatParams(netD, lr)
        self.beta = float(beta)
        self._attach()

    n_rands = 3  # noise, eps of the decoder phase, eps of the encoder phase

    def flat_params(self):
        return [self.feg, self.fd]

    def _deferred(self):
        return [self.feg] if (self.DEFER and not self.feg.shard) else []

    def step(self, data, real_label=None, fake_label=None, noise=None, eps_dec=None, eps_enc=None):
        if self._graph is not None:
            rands = None if (noise is None and eps_dec is None and eps_enc is None) else [noise, eps_dec, eps_enc]
            return self._graph_step(data, real_label, fake_label, rands)
        if real_label is None:
            real_label, fake_label = self.draw_labels()
        return self._step_impl(data, real_label, fake_label, noise, eps_dec, eps_enc)

    # DM_STACK_DEC=1 (default): the encoder / decoder forward of the "decoder" phase is computed at the START of the step.
    # It reads only the encoder / decoder parameters, which the discriminator phase does not change, so decode(noise)
    # (discriminator phase) and decode(z) (decoder phase) go through every decoder GEMM TOGETHER as one stacked
    # two-group pass (BatchNorm per group, noise first: the reference's order of running-stat updates), and so do their
    # two backward passes.  =0: the reference's literal order, two separate decoder passes each way.
    STACK_DEC = os.environ.get("DM_STACK_DEC", "1") != "0"

    def _step_impl(self, data, real_label, fake_label, noise=None, eps_dec=None, eps_enc=None):
        feg, fd, dev = self.feg, self.fd, data.device
        b = data.shape[0]
        (errD_real, errD_fake, sum_dx, errG_fake, errG_recon, sim_loss, loss_dec, kld, loss_enc) = _scalars(9, dev)
        # every image the networks read, as padded bf16 images stacked [data | fake | recon]: each producer writes its
        # slice once, D's stacked passes and the two encoder forwards read them in place (no torch.cat, no im2col)
        # sharded Adam (data parallel): the encoder-phase update of the previous step reaches the other ranks now, under
        # the discriminator phase, which does not read the encoder's weights
        feg.gather_if_pending()
        if self.DEFER:  # single GPU: the same for the Adam update of the encoder's two 33.5 M-element Linear weights
            feg.finish_big(self._side())
        pim = ops.pim_empty(3 * b, dev)
        data = self._ingest(data, pim[:b])
        nt = b * self.dist.world
        if noise is None:
            noise = torch.randn(b, 128, device=dev)
        stack = self.STACK_DEC

        if stack:
            # ---- encoder forward + BOTH decoder forwards of the discriminator / decoder phases (:97 fake, :127 recon)
            feg.zero_grad()
            mu, logvar, Se = engine.encoder_forward(None, feg.P, feg.buffers, feg.cache, True, pim=pim[:b],
                                                    before_heads=self._before_use(feg))
            if eps_dec is None:
                eps_dec = torch.randn_like(mu)
            code = torch.empty((2 * b, 128), dtype=BF16, device=dev)
            ops.cast_bf16(noise.contiguous(), code[:b])
            ops.reparam_forward(mu, logvar, eps_dec, out_bf16=code[b:])
            both, Sg12 = engine.decoder_forward(code, feg.P, feg.buffers, feg.cache, True, pim_out=pim[b:], groups=2)
            recon = both[b:]

        # ================= discriminator phase (:95-123).  D(data) and D(fake.detach()) share every GEMM launch
        # (stacked along the batch); BatchNorm statistics / running-stat updates stay per pass, real first.
        fd.zero_grad()
        if not stack:
            fake, Sg1 = engine.decoder_forward(noise, feg.P, feg.buffers, feg.cache, True, pim_out=pim[b:2 * b])
        prob, _, S12 = engine.discriminator_forward(None, fd.P, fd.buffers, fd.cache, True, groups=2, pim=pim[:2 * b])
        dprob = torch.empty_like(prob)
        ops.bce_const(prob[:b], real_label, errD_real, 1.0, n_total=nt, dprob=dprob[:b], stat=sum_dx)
        ops.bce_const(prob[b:], fake_label, errD_fake, 1.0, n_total=nt, dprob=dprob[b:])
        # (stacked order: nothing independent follows the discriminator's backward pass, so the update of its
        # 33.5 M-element Linear weight starts on the side stream as soon as that layer is back-propagated)
        engine.discriminator_backward(S12, dprob, None, fd.P, fd.G, fd.cache, False, True, overwrite_big=True,
                                      grad_ready=self._early(fd), linear_done=self._heads_done(fd) if stack else None)
        del S12
        if stack:
            fd.reduce_rest_and_wait(self.dist)  # data parallel: D's gradient all-reduce runs on the NCCL stream
            fd.adam()
        else:
            fd.reduce_rest(self.dist)
            # ... and D's Adam update (HBM-bound) runs on a side stream, both under the tensor-bound encoder / decoder
            # forward below
            fork = self._fork_side(lambda: (self.dist.wait(), fd.adam()))

        # ================= "decoder" phase (:127-164): gradient of
        #   BCE(D(fake), real) + BCE(D(recon), real) + 0.5*||Dis_l(recon) - Dis_l(x)||^2 + ||recon - x||^2
        # w.r.t. ALL encoder and decoder parameters, D frozen at its updated value
        if not stack:
            # ... while the encoder / decoder forward of this phase, which does not read D, is computed; the D update
            # (:123) lands before D is evaluated again, as in the reference
            feg.zero_grad()
            mu, logvar, Se = engine.encoder_forward(None, feg.P, feg.buffers, feg.cache, True, pim=pim[:b],
                                                    before_heads=self._before_use(feg))
            if eps_dec is None:
                eps_dec = torch.randn_like(mu)
            _, z16 = ops.reparam_forward(mu, logvar, eps_dec)
            recon, Sg2 = engine.decoder_forward(z16, feg.P, feg.buffers, feg.cache, True, pim_out=pim[2 * b:])
            self._join_side(fork)
        # D(data) | D(fake) | D(recon) in one stacked pass (BatchNorm per pass, in the reference's order :129,147,150)
        prob3, feat3, S345 = engine.discriminator_forward(None, fd.P, fd.buffers, fd.cache, True, groups=3, pim=pim,
                                                          before_linear=self._before_use(fd))
        sim_real, sim_recon = feat3[:b], feat3[2 * b:]
        dprob2 = torch.empty(2 * b, dtype=F32, device=dev)
        ops.bce_const(prob3[b:2 * b], real_label, errG_fake, 1.0, n_total=nt, dprob=dprob2[:b])
        ops.bce_const(prob3[2 * b:], real_label, errG_recon, 1.0, n_total=nt, dprob=dprob2[b:])
        dfeat2 = torch.zeros((2 * b, 2048), dtype=F32, device=dev)
        ops.mse_sum(sim_recon, sim_real, sim_loss, 0.5, dfeat2[b:], 0.5)
        # back through D for the fake and recon passes only: sim_real's path into D (and every D weight gradient of
        # this phase) is discarded by the reference's next netD.zero_grad() (:95)
        dx = engine.discriminator_backward(S345, dprob2, dfeat2, fd.P, None, fd.cache, True, False, group_range=(1, 3))
        del S345
        dfake, drecon = dx[:b], dx[b:]
        if stack:
            ops.mse_sum(recon, data, loss_dec, 1.0, drecon, 1.0, accumulate=True)
            dz = engine.decoder_backward(Sg12, dx, feg.P, feg.G, feg.cache, True, True, overwrite_big=True)[b:]
            del Sg12
        else:
            engine.decoder_backward(Sg1, dfake, feg.P, feg.G, feg.cache, False, True, overwrite_big=True)
            del Sg1
            ops.mse_sum(recon, data, loss_dec, 1.0, drecon, 1.0, accumulate=True)
            dz = engine.decoder_backward(Sg2, drecon, feg.P, feg.G, feg.cache, True, True)
            del Sg2
        feg.reduce_from(self.dist, "preprocess.0.weight")  # decoder gradients are final: reduce them under the encoder backward
        _, _, dmu, dlv = ops.reparam_backward(dz, logvar, eps_dec)
        engine.encoder_backward(Se, dmu, dlv, feg.P, feg.G, feg.cache, True, overwrite_big=True,
                                grad_ready=self._early(feg), heads_done=self._heads_done(feg))
        del Se
        feg.reduce_rest_and_wait(self.dist)
        # (big Linear weights on the side stream: the encoder convolutions of the next phase do not read them)
        feg.adam(big="side" if self.DEFER else "now", side=self._side())

        # ================= "encoder" phase (:167-193): gradient of beta*KL + ||recon - x||^2, fresh forward
        feg.zero_grad()
        mu, logvar, Se = engine.encoder_forward(None, feg.P, feg.buffers, feg.cache, True, pim=pim[:b],
                                                before_heads=self._before_use(feg))
        if eps_enc is None:
            eps_enc = torch.randn_like(mu)
        _, z16 = ops.reparam_forward(mu, logvar, eps_enc)
        recon, Sg3 = engine.decoder_forward(z16, feg.P, feg.buffers, feg.cache, True)
        drecon = torch.empty_like(recon)
        ops.mse_sum(recon, data, loss_enc, 1.0, drecon, 1.0)
        dmu_kl, dlv_kl = torch.empty_like(mu), torch.empty_like(mu)
        ops.kl(mu, logvar, kld, self.beta, dmu_kl, dlv_kl)
        dz = engine.decoder_backward(Sg3, drecon, feg.P, feg.G, feg.cache, True, True, overwrite_big=True)
        feg.reduce_from(self.dist, "preprocess.0.weight")
        _, _, dmu, dlv = ops.reparam_backward(dz, logvar, eps_enc, dmu_kl, dlv_kl)
        engine.encoder_backward(Se, dmu, dlv, feg.P, feg.G, feg.cache, True, overwrite_big=True,
                                grad_ready=self._early(feg), heads_done=self._heads_done(feg))
        feg.reduce_rest_and_wait(self.dist)
        # (sharded: its all-gather rides under the next step's discriminator phase; single GPU: so does the update of the
        # two big Linear weights itself)
        feg.adam(gather=False, big="defer" if self.DEFER else "now")
        self.metrics = {"errD_real": errD_real, "errD_fake": errD_fake, "D_x": sum_dx / b, "errG_fake": errG_fake,
                        "errG_recon": errG_recon, "sim": sim_loss, "recon_dec": loss_dec, "kld": kld,
                        "recon_enc": loss_enc}
        return self.metrics

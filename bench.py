#!/usr/bin/env python
"""Benchmark of the sequence-loss hot path (BASELINE.json metric: RNN-T/CTC loss fwd+bwd utterances/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rnnt_cfg3|rnnt_cfg4|ctc_cfg2|ctc_cfg1] [--precision bf16|fp32]
                    [--lengths full|ragged] [--grad-payload-mb M] [--no-extras]

A "step" is one pass of the hot path over one batch of synthetic input:
  RNN-T  enc_proj = w_enc(eouts), dec_proj = w_dec(douts) -> fused joint + log-softmax + transducer loss ->
         backward to eouts, douts and every joint parameter (asr/modeling/decoders/rnn_transducer.py:57-58,101-115,
         147-156); by default the projections and their backward run inside the library too (two C calls + the
         lattice per step), --unfolded keeps them as cuBLAS Linear around the fused joint
  CTC    eouts -> output Linear -> log-softmax + CTC loss -> backward to eouts and the Linear's parameters
         (asr/modeling/decoders/ctc.py:103-113): the fused tensor-core head from 8192 frames per batch upwards,
         else cuBLAS Linear + the fused loss kernels on logits
Default workload (N=1): BASELINE cfg 3, "RNN-T(Cf.) 1kBPE 26M fused joint+loss, B=32 T=250 U=100 V=1024"; nothing
of size N x V is written to HBM, forward or backward.  The default N=1 run also reports, under "extra", CTC cfg 2,
and under "gpu_baseline" same-box GPU comparators (the reference's materialised op sequence on CUDA; torch's CUDA
ctc_loss) -- `--no-extras` skips them.
For N>1 every rank processes its own batch of the same shape (weak scaling, batch-sharded) and the gradients of
the path's parameters (plus, with --grad-payload-mb, a stand-in for the rest of the model's gradients) are
all-reduced over NCCL inside the timed step, launched from autograd hooks (the fused joint's six parameter gradients
arrive in one flat buffer: one collective).
The end-to-end number is the median of three regions of `--steps` steps fed from pinned host buffers; EMO_BENCH_TRACE=1
prints every timed step's device time to stderr.

One JSON line is printed by rank 0 (see the driver contract in the task statement).
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, B, T, U, V, He, Hd, J
    "rnnt_cfg3": dict(kind="rnnt", B=32, T=250, U=100, V=1024, He=256, Hd=512, J=512,
                      desc="RNN-T(Cf.) 1kBPE 26M (L4-style) fused joint+loss, B=32 T=250 U=100 V=1024"),
    "rnnt_cfg4": dict(kind="rnnt", B=8, T=1000, U=400, V=4096, He=512, Hd=512, J=512,
                      desc="RNN-T(Cf.) 4kBPE Large 91M (C6-style), V=4096 joint, T=1000 U=400, B=8"),
    "ctc_cfg2": dict(kind="ctc", B=64, T=374, U=80, V=5000, He=256,
                     desc="CTC(Cf.) 23M (L2-style), batch 64, ~15 s utterances (T=374), BPE vocab 5k"),
    "ctc_cfg1": dict(kind="ctc", B=8, T=249, U=60, V=10872, He=256,
                     desc="CTC(Trf.) 20M (L1-style) CTC loss, batch 8, T=249, V=10872"),
}
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "dram_traffic.json")   # written by tools/ncu_traffic.py from ncu --set full


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained"),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def load_traffic(workload, route="ring"):
    """dram__bytes_read.sum + dram__bytes_write.sum per call, from the committed ncu --set full capture of this
    workload (profiles/dram_traffic.json names the capture it came from).  None if there is no capture."""
    try:
        d = json.load(open(TRAFFIC_FILE))
        e = d[workload][route]
        return e, d.get("_source")
    except (OSError, KeyError, ValueError):
        return None, None


# ------------------------------------------------------------------------------------------------
def make_lengths(B, T, U, regime, gen):
    if regime == "full":
        return torch.full((B,), T, dtype=torch.long), torch.full((B,), U, dtype=torch.long)
    r = torch.sort(torch.rand(B, generator=gen) * 0.4 + 0.6, descending=True).values
    r[0] = 1.0
    return (T * r).long().clamp(min=1), (U * r).long()


def make_labels(B, U, V, gen):
    # labels in {1} u [4,V): specials per corpora/utils/spm_train.py:7-9; padding = eos (2)
    pool = torch.cat([torch.tensor([1]), torch.arange(4, V)])
    return pool[torch.randint(len(pool), (B, U), generator=gen)]


class RNNTWorkload:
    """Synthetic cfg-3/cfg-4 joint inputs: eouts ~ N(0,1) (encoder ends in LayerNorm), douts =
    tanh(N(0,1)) (LSTM range), nn.Linear default init."""

    def __init__(self, w, seed, regime):
        gen = torch.Generator().manual_seed(seed)
        B, T, U, V = w["B"], w["T"], w["U"], w["V"]
        self.w = w
        self.eouts = torch.randn(B, T, w["He"], generator=gen)
        self.douts = torch.tanh(torch.randn(B, U + 1, w["Hd"], generator=gen))
        self.tlen, self.ulen = make_lengths(B, T, U, regime, gen)
        self.ys = make_labels(B, U, V, gen)
        for b in range(B):
            self.ys[b, self.ulen[b]:] = 2
        torch.manual_seed(1234)  # identical parameters on every rank
        self.w_enc = torch.nn.Linear(w["He"], w["J"])
        self.w_dec = torch.nn.Linear(w["Hd"], w["J"])
        self.output = torch.nn.Linear(w["J"], V)
        self.n_valid = int((self.tlen * (self.ulen + 1)).sum())

    def algorithmic_flops(self):
        w = self.w
        return 6.0 * self.n_valid * w["J"] * w["V"] + 6.0 * w["B"] * (w["T"] * w["He"] + (w["U"] + 1) * w["Hd"]) * w["J"]

    def joint_gemm_flops(self):
        return 2.0 * self.n_valid * self.w["J"] * self.w["V"]


class CTCWorkload:
    """Synthetic cfg-1/cfg-2 CTC head inputs: eouts ~ N(0,1) (B,T,He), output Linear(He,V) default init."""

    def __init__(self, w, seed, regime):
        gen = torch.Generator().manual_seed(seed)
        B, T, U, V = w["B"], w["T"], w["U"], w["V"]
        self.w = w
        self.eouts = torch.randn(B, T, w["He"], generator=gen)
        self.tlen, _ = make_lengths(B, T, U, regime, gen)
        self.ulen = torch.randint(U // 2, U + 1, (B,), generator=gen)
        self.ys = make_labels(B, U, V, gen)
        torch.manual_seed(1234)
        self.output = torch.nn.Linear(w["He"], V)


def run_e2e(step, host, dev, steps, first=lambda out: out):
    """`steps` steps through the public API with HOST inputs: every step's inputs are copied from pinned host
    memory (on a copy stream, issued while the previous step computes -- what a prefetching loader does) and
    every step's loss is copied back to pinned host memory and read by the host.  The read of step i's loss
    happens after step i+1 has been enqueued (one step late, as a logging loop does), so the device never waits
    for the host.  Returns wall-clock seconds for exactly `steps` steps, all losses read."""
    key = (dev.index, "copy")
    if key not in _E2E_STATE:   # one copy stream / one pair of result buffers per process: the allocator's pool of
        _E2E_STATE[key] = (torch.cuda.Stream(device=dev),                      # upload blocks is per stream
                           [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)])
    copy_stream, pinned = _E2E_STATE[key]
    main = torch.cuda.current_stream(dev)

    def upload():
        with torch.cuda.stream(copy_stream):
            tens = [h.to(dev, non_blocking=True) for h in host]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return tens, ev

    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    nxt = upload()
    pending = None   # (pinned scalar, event) of the previous step
    total = 0.0
    for i in range(steps):
        cur, ev = nxt
        main.wait_event(ev)
        for t in cur:
            t.record_stream(main)
        out = step(*cur)
        buf = pinned[i & 1]
        buf.copy_(first(out).detach().float(), non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        if i + 1 < steps:
            nxt = upload()
        if pending is not None:
            pending[1].synchronize()
            total += float(pending[0])
        pending = (buf, done)
    pending[1].synchronize()
    total += float(pending[0])
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t0


_E2E_STATE = {}


def e2e_regions(step, host, dev, steps, barrier, reps=3):
    """The end-to-end number: 2 untimed steps, then `reps` regions of exactly `steps` steps each; returns the MEDIAN
    region's seconds and all of them.  Long-lived Python objects are moved out of the garbage collector's reach
    first (gc.freeze, what a training loop does once its model is built): a full collection of a process with
    torch loaded takes ~100 ms, more than a whole 20-step region, and one landing inside it was measured to
    triple the per-step time (tools/diag_e2e.py)."""
    run_e2e(step, host, dev, 2)
    gc.collect()
    gc.freeze()
    secs = []
    for _ in range(reps):
        barrier()
        secs.append(run_e2e(step, host, dev, steps))
    barrier()
    return statistics.median(secs), secs


def timed_steps(step, resident, steps, flush, barrier):
    """`steps` steps with device-resident inputs: per-step CUDA events, L2 flushed (untimed) between steps."""
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(*resident)
        b.record()
        evs.append((a, b))
    barrier()
    ms = [a.elapsed_time(b) for a, b in evs]
    LAST_STEP_MS[:] = ms
    if os.environ.get("EMO_BENCH_TRACE"):
        sys.stderr.write("rank %s step ms: %s\n" % (os.environ.get("RANK", "0"), " ".join("%.3f" % x for x in ms)))
    return sum(ms)


LAST_STEP_MS = []   # per-step device times of the most recent timed_steps() call (this rank)


def step_spread(ms):
    """first / median / last-quarter mean of the per-step device times: the GPU leaves its burst clocks after
    ~0.1 s of sustained tensor load (power cap), so late steps of a long region run several per cent slower."""
    q = max(1, len(ms) // 4)
    return {"first": round(ms[0], 4), "median": round(statistics.median(ms), 4), "min": round(min(ms), 4),
            "max": round(max(ms), 4), "first_quarter_mean": round(statistics.mean(ms[:q]), 4),
            "last_quarter_mean": round(statistics.mean(ms[-q:]), 4)}


def time_call(fn, flush, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    kev = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        kev.append((a, b))
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in kev)


# ------------------------------------------------------------------------------------------------
def rnnt_step_fn(E, wl, precision, params, reducer, payload, world, folded=None):
    """folded (default: whenever the tensor-core mode takes the shape): the w_enc / w_dec projections and their backward
    run inside the library (emo_rnnt_joint_full_*); otherwise plain cuBLAS Linear around the fused joint."""
    crit = E.RNNTJointLoss(blank_id=0, precision=precision)
    w = wl.w
    if folded is None:
        from emoasr_b200 import functional as EF
        folded = precision == "bf16" and EF.joint_full_supported(w["B"], w["T"], w["U"] + 1, w["He"], w["Hd"], w["J"], w["V"])
    full = E.RNNTJointFullLoss(blank_id=0)

    def step(eouts, douts, ys, tlen, ulen):
        for p in params:
            p.grad = None           # optimizer.zero_grad(): autograd assigns the new gradients
        eouts = eouts.detach().requires_grad_()
        douts = douts.detach().requires_grad_()
        if folded:
            loss = full(eouts, douts, wl.w_enc, wl.w_dec, wl.output, ys, tlen, ulen)
        else:
            loss = crit(wl.w_enc(eouts), wl.w_dec(douts), wl.output.weight, wl.output.bias, ys, tlen, ulen)
        loss.backward()             # every gradient's all-reduce is launched from an autograd hook during this call
        if world > 1:
            if payload is not None:
                reducer.start_extra(payload)   # stand-in for the rest of the model's gradients
            reducer.finish()        # compute stream waits for the collectives (no host sync)
        return loss
    return step


def run_ours_rnnt(args, w, rank, world, dev):
    import emoasr_b200 as E
    from emoasr_b200 import _lib, sharding

    # tensor-core mode: the two small projection GEMMs around the fused op (plain cuBLAS Linear, as in the
    # reference's joint) run in TF32 instead of SIMT fp32; fp32 mode keeps them in true fp32
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "bf16"
    wl = RNNTWorkload(w, seed=rank, regime=args.lengths)
    mods = torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
    params = [wl.output.weight, wl.output.bias, wl.w_dec.weight, wl.w_dec.bias, wl.w_enc.weight, wl.w_enc.bias]
    buckets, payload = None, None
    if world > 1:
        buckets = sharding.GradReducer(params)   # per-gradient all-reduce from autograd hooks, in place
        if args.grad_payload_mb > 0:
            payload = torch.zeros(int(args.grad_payload_mb * 1e6) // 4, device=dev)
    host = [t.pin_memory() for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    step = rnnt_step_fn(E, wl, args.precision, params, buckets, payload, world, folded=False if args.unfolded else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(*resident)
    gc.collect()
    gc.freeze()     # see e2e_regions: no full collection of torch's long-lived objects inside a timed region
    # Everything rank-specific happens BEFORE the barrier that opens the timed region.  Rank 0 alone starts the clock
    # sampler (NVML initialisation: ~10 ms on the host); started after the barrier, that delay sat inside every other
    # rank's first timed step (they wait for rank 0 in the first collective: 13.5 ms instead of 3.3 -- measured with
    # EMO_BENCH_TRACE=1), i.e. +0.5 ms per step on a 20-step region, the larger part of the "scaling loss" of
    # earlier rounds.
    mon = ClockMonitor(dev.index if dev.index is not None else 0, enabled=rank == 0)
    mon.start()
    barrier()
    ms_total = timed_steps(step, resident, args.steps, flush, barrier)
    clocks = mon.stop()
    spread = step_spread(LAST_STEP_MS)
    # ---- end-to-end: host (pinned) buffers in, loss value out, wall clock
    e2e_s, e2e_all = e2e_regions(step, host, dev, args.steps, barrier)

    prec = 1 if args.precision == "bf16" else 0
    B, T, U1, J, V = w["B"], w["T"], w["U"] + 1, w["J"], w["V"]
    from emoasr_b200 import functional as EF
    folded = prec == 1 and EF.joint_full_supported(B, T, U1, w["He"], w["Hd"], J, V) and not args.unfolded
    per_step = (_lib.launch_count(_lib.OP_RNNT_JOINT_FULL, prec, B, T, U1, J, V) if folded else
                _lib.launch_count(_lib.OP_RNNT_JOINT_FWD, prec, B, T, U1, J, V)
                + _lib.launch_count(_lib.OP_RNNT_JOINT_BWD, prec, B, T, U1, J, V))
    out = dict(ms_total=ms_total, e2e_s=e2e_s, e2e_all=e2e_all, units=w["B"], clocks=clocks, roofline=None, launches=per_step * args.steps,
               h2d=sum(t.numel() * t.element_size() for t in host), d2h=4, flops=wl.algorithmic_flops(),
               n_valid=wl.n_valid, extra={"step_ms_rank0": spread})
    if world > 1:
        # diagnosis of the scaling loss: every rank's step WITHOUT the collective (same kernels, same inputs).  The
        # synchronised step can never be faster than the slowest GPU's local step.
        local = rnnt_step_fn(E, wl, args.precision, params, None, None, 1, folded=False if args.unfolded else None)
        buckets.enabled = False
        for _ in range(3):
            local(*resident)
        n = max(5, args.steps // 2)
        lms = timed_steps(local, resident, n, flush, barrier) / n
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = lms
        dist.all_reduce(t)
        out["extra"]["per_rank_local_step_ms"] = [round(float(x), 4) for x in t]
    if args.precision == "bf16" and rank == 0:
        out["roofline"] = rnnt_kernel_roofline(args, w, wl, resident, flush, dev)
    return out


def rnnt_kernel_roofline(args, w, wl, resident, flush, dev):
    """The dominant kernels alone, through the C ABI, timed with CUDA events: emo_rnnt_joint_bwd (ring route: one
    persistent kernel executing 3 GEMM units for 2 algorithmic ones + casts, ring prep, axis reductions) and
    emo_rnnt_joint_fwd."""
    import ctypes
    from emoasr_b200 import _lib
    lib = _lib.load()
    with torch.no_grad():
        enc_proj = wl.w_enc(resident[0]).contiguous()
        dec_proj = wl.w_dec(resident[1]).contiguous()
    B, T, J = enc_proj.shape
    U1, V = dec_proj.size(1), w["V"]
    ws = torch.empty(_lib.workspace_bytes(_lib.OP_RNNT_JOINT_FWD, 1, B, T, U1, J, V), dtype=torch.uint8, device=dev)
    lp2 = torch.empty(B, T, U1, 2, device=dev)
    lse = torch.empty(B, T, U1, device=dev)
    wo, bo = wl.output.weight.detach().contiguous(), wl.output.bias.detach().contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        rc = lib.emo_rnnt_joint_fwd(p(enc_proj), p(dec_proj), p(wo), p(bo), p(resident[2]), p(resident[3]),
                                    p(resident[4]), B, T, U1, J, V, 0, 1, p(lp2), p(lse), p(ws), ws.numel(), st)
        _lib.check(rc, "emo_rnnt_joint_fwd")
    alpha = torch.empty(B, T, U1, device=dev); beta = torch.empty(B, T, U1, device=dev)
    cost = torch.empty(B, device=dev); gamma2 = torch.empty(B, T, U1, 2, device=dev)
    call()
    _lib.check(lib.emo_rnnt_lattice_fwd_bwd(p(lp2), p(resident[3]), p(resident[4]), B, T, U1, p(alpha), p(beta),
                                            p(cost), p(gamma2), st), "emo_rnnt_lattice_fwd_bwd")
    gcost = torch.full((B,), 1.0 / B, device=dev)
    wsb = torch.empty(_lib.workspace_bytes(_lib.OP_RNNT_JOINT_BWD, 1, B, T, U1, J, V), dtype=torch.uint8, device=dev)
    d_enc, d_dec = torch.empty_like(enc_proj), torch.empty_like(dec_proj)
    d_w, d_b = torch.empty_like(wo), torch.empty_like(bo)

    def call_bwd():
        rc = lib.emo_rnnt_joint_bwd(p(enc_proj), p(dec_proj), p(wo), p(bo), p(resident[2]), p(resident[3]),
                                    p(resident[4]), p(lse), p(lp2), p(gamma2), p(gcost), p(None), B, T, U1, J, V,
                                    0, 1, p(d_enc), p(d_dec), p(d_w), p(d_b), p(wsb), wsb.numel(), st)
        _lib.check(rc, "emo_rnnt_joint_bwd")

    # Kernel-alone timing against the BURST peak: by now the GPU has run ~100 back-to-back steps and sits on its power
    # cap (steps 3.2 -> 3.6 ms, extra.step_ms_rank0 / profiles/r2g_multi_gpu_step_traces.txt); MEASURED_PEAKS' burst
    # figure is a short cuBLAS run on an idle GPU, so each call is timed the same way: after one second of idle.
    iters = max(args.steps, 5)
    torch.cuda.synchronize()
    time.sleep(1.0)
    f_ms = time_call(call, flush, iters)
    torch.cuda.synchronize()
    time.sleep(1.0)
    b_ms = time_call(call_bwd, flush, iters)
    peaks = load_peaks()
    unit = wl.joint_gemm_flops()

    def roofline(kernel, flops, ms, extra):
        ach = flops / (ms * 1e-3) / 1e12
        r = {"bound": "tensor", "kernel": kernel, "achieved": round(ach, 1), "peak": peaks["tf_burst"],
             "unit": "TFLOP/s", "frac": round(ach / peaks["tf_burst"], 4), "traffic": None,
             "kernel_ms": round(ms, 4), "peak_source": peaks["src"] + ", burst",
             "timing": f"the C-ABI call alone: 1 s idle, 3 warm-up + {iters} timed launches, CUDA events on the launching "
                       "stream, L2 flushed between launches",
             "algorithmic_flops_per_launch": flops}
        r.update(extra)
        return r
    roof = roofline("emo_rnnt_joint_bwd = joint_bwd_ring_kernel (+ casts, ring prep, axis reductions)",
                    2 * unit, b_ms,
                    {"executed_flops_per_launch": 3 * unit,
                     "executed_frac": round(3 * unit / (b_ms * 1e-3) / 1e12 / peaks["tf_burst"], 4),
                     "note": "algorithmic = dh and dW GEMMs (2 x 2*N*J*V); the kernel also recomputes the logits "
                             "(3 GEMM units executed) and hands dz to the two gradient GEMMs through an "
                             "L2-resident ring: nothing of size N x V reaches HBM"})
    roof["forward"] = roofline("emo_rnnt_joint_fwd = joint_fwd_kernel (+ weight / stream casts)", unit, f_ms, {})
    tr, src = load_traffic(args.workload)
    if tr is not None and args.lengths == "full":
        roof["traffic"] = tr.get("bwd")
        roof["forward"]["traffic"] = tr.get("fwd")
        roof["traffic_step"] = tr.get("step")
        roof["traffic_source"] = src
    return roof


# ------------------------------------------------------------------------------------------------
def ctc_step_fn(E, wl, B, fused=True):
    """fused: Linear + log_softmax + CTC + the Linear's backward as one tensor-core op per direction (no (B,T,V)
    tensor in memory); unfused: cuBLAS Linear (TF32) around the fp32 loss kernels on raw logits."""
    crit = E.CTCHeadLoss(blank=0)

    def step(eouts, ys, tlen, ulen):
        wl.output.weight.grad = None
        wl.output.bias.grad = None
        eouts = eouts.detach().requires_grad_()
        if fused:
            loss = crit(eouts, wl.output.weight, wl.output.bias, ys, tlen, ulen)      # ctc.py:103-113
        else:
            logits = wl.output(eouts)                                  # ctc.py:103
            loss = E.ctc_loss(logits, ys, tlen, ulen, blank=0, reduction="sum") / B   # ctc.py:109-113
        loss.backward()
        return loss
    return step


def run_ours_ctc(args, w, rank, world, dev):
    import emoasr_b200 as E
    from emoasr_b200 import _lib

    # the head's Linear(He,V) and its two backward GEMMs via cuBLAS in TF32 (as the RNN-T leg's projections); in
    # true fp32 (SIMT) they alone take ~4 ms at cfg 2, ten times the loss kernels
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    wl = CTCWorkload(w, seed=rank, regime=args.lengths)
    wl.output.to(dev)
    B, T, V, He = w["B"], w["T"], w["V"], w["He"]
    host = [t.pin_memory() for t in (wl.eouts, wl.ys, wl.tlen, wl.ulen)]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    from emoasr_b200 import functional as EF
    # the drop-in decoder's rule (decoders.FusedCTCForward): fused head from 8192 frames per batch upwards
    fused = (EF.ctc_head_supported(B, T, He, V, w["U"]) and not getattr(args, "ctc_unfused", False)
             and (B * T >= 8192 or getattr(args, "ctc_fused", False)))
    step = ctc_step_fn(E, wl, B, fused)
    for _ in range(args.warmup):
        step(*resident)
    torch.cuda.synchronize()
    gc.collect()
    gc.freeze()
    mon = ClockMonitor(dev.index if dev.index is not None else 0, enabled=rank == 0)
    mon.start()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    barrier()
    ms_total = timed_steps(step, resident, args.steps, flush, barrier)
    clocks = mon.stop()
    e2e_s, e2e_all = e2e_regions(step, host, dev, args.steps, barrier)
    peaks = load_peaks()
    # SURVEY 8d's per-step figure for the CTC path is the logits traffic of the unfused sequence, 3*B*T*V*4 bytes
    # (read twice, gradient written once).  The fused head moves none of it: what bounds it is the tensor pipe on
    # the head's three GEMMs (Linear forward, d_eouts, d_W: 6*B*T*He*V flop).  Both fractions are reported: the
    # HBM one as "time the unfused byte stream would need at peak / step time" (> 1 is possible when fused).
    n_frames = int(wl.tlen.sum())
    alg_bytes = 3.0 * B * T * V * 4
    ms = ms_total / args.steps
    ach = alg_bytes / (ms * 1e-3) / 1e9
    if fused:
        flops = 6.0 * n_frames * He * V
        tf = flops / (ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "fused CTC head step: joint_fwd_kernel<plain> + emissions + lattices + "
                                             "joint_bwd_ring_kernel (plain) + sparse label kernels",
                "achieved": round(tf, 1), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": round(tf / peaks["tf_burst"], 4), "traffic": None, "peak_source": peaks["src"] + ", burst",
                "algorithmic_flops_per_step": flops,
                "hbm_equivalent": {"achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                                   "frac": round(ach / peaks["hbm"], 4), "algorithmic_bytes_per_step": alg_bytes},
                "note": "small-K GEMMs (K = He = 256) with an exp per logit: the step is bound by the lattice "
                        "recursions' latency and the MUFU epilogues, not by the tensor pipe or HBM"}
    else:
        roof = {"bound": "hbm", "kernel": "CTC head step: output Linear fwd/bwd (cuBLAS TF32) + row_lse + lattices + grad kernels",
                "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4),
                "traffic": None, "peak_source": peaks["src"], "algorithmic_bytes_per_step": alg_bytes,
                "note": "algorithmic bytes = 3*B*T*V*4 (the loss kernels' logits traffic); the step timed here also "
                        "contains the Linear(He,V) forward and both of its backward GEMMs (cuBLAS, TF32)"}
    roof["head"] = "fused" if fused else "unfused"
    tr, src = load_traffic(args.workload, "fused" if fused else "default")
    if tr is not None:
        roof["traffic"], roof["traffic_source"] = tr.get("step"), src
    if fused and rank == 0:
        ustep = ctc_step_fn(E, wl, B, False)
        for _ in range(3):
            ustep(*resident)
        torch.cuda.synchronize()
        n = max(5, args.steps // 2)
        roof["unfused_ms_per_step"] = round(timed_steps(ustep, resident, n, flush, torch.cuda.synchronize) / n, 4)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    return dict(ms_total=ms_total, e2e_s=e2e_s, e2e_all=e2e_all, units=B, clocks=clocks, roofline=roof,
                launches=(_lib.launch_count(_lib.OP_CTC_HEAD, 1, B, T, 1, He, V) if fused
                          else _lib.launch_count(_lib.OP_CTC, 0, B, T, 1, 1, V)) * args.steps,
                h2d=sum(t.numel() * t.element_size() for t in host), d2h=4, flops=None, n_valid=None, extra={})


# ------------------------------------------------------------------------------------------------
def gpu_baselines(dev, steps=3):
    """Same-box GPU comparators (rank 0, N=1): what the REFERENCE's op sequence costs on this B200.
      rnnt_cfg3  joint -> log_softmax -> transducer loss with all three (B,T,U+1,V) fp32 tensors materialised
                 (rnn_transducer.py:101-115); warp_rnnt is absent from the image, torchaudio's CUDA rnnt_loss
                 (fused_log_softmax=False) has the same contract.
      ctc_cfg2   output Linear -> transpose -> log_softmax -> torch's CUDA ctc_loss (ctc.py:103-113)."""
    out = {}
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False     # the reference runs cuBLAS in fp32
    # ---- CTC cfg 2
    try:
        w = WORKLOADS["ctc_cfg2"]
        wl = CTCWorkload(w, 0, "full")
        wl.output.to(dev)
        e, ys, tl, ul = (t.to(dev) for t in (wl.eouts, wl.ys, wl.tlen, wl.ulen))
        fn = torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=True)

        def ctc_step():
            wl.output.zero_grad(set_to_none=True)
            x = e.detach().requires_grad_()
            logits = wl.output(x)
            loss = fn(logits.transpose(1, 0).log_softmax(dim=2), ys, tl, ul) / logits.size(0)
            loss.backward()
        for _ in range(2):
            ctc_step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            ctc_step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out["ctc_cfg2"] = {"value": round(w["B"] / (ms * 1e-3), 1), "unit": "utt/s", "ms_per_step": round(ms, 3),
                           "what": "torch CUDA: Linear -> log_softmax -> nn.CTCLoss -> backward (the reference's ops)"}
    except Exception as ex:   # noqa: BLE001
        out["ctc_cfg2"] = {"unavailable": repr(ex)[:200]}
    torch.cuda.empty_cache()
    # ---- RNN-T cfg 3, materialised
    try:
        import torchaudio
        w = WORKLOADS["rnnt_cfg3"]
        wl = RNNTWorkload(w, 0, "full")
        mods = torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
        e, d, ys, tl, ul = (t.to(dev) for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int()))

        def rnnt_step():
            mods.zero_grad(set_to_none=True)
            x, y = e.detach().requires_grad_(), d.detach().requires_grad_()
            logits = wl.output(torch.tanh(wl.w_enc(x).unsqueeze(2) + wl.w_dec(y).unsqueeze(1)))
            lp = torch.log_softmax(logits, dim=-1)
            loss = torchaudio.functional.rnnt_loss(lp, ys, tl, ul, blank=0, reduction="mean", fused_log_softmax=False)
            loss.backward()
        rnnt_step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            rnnt_step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out["rnnt_cfg3"] = {"value": round(w["B"] / (ms * 1e-3), 1), "unit": "utt/s", "ms_per_step": round(ms, 2),
                            "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1),
                            "what": "torch CUDA: joint (cuBLAS fp32) -> log_softmax -> torchaudio rnnt_loss "
                                    "(fused_log_softmax=False, warp_rnnt's contract) -> backward, B=32, all "
                                    "(B,T,U+1,V) tensors materialised"}
    except Exception as ex:   # noqa: BLE001
        out["rnnt_cfg3"] = {"unavailable": repr(ex)[:200]}
    torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = tf32
    return out


# ------------------------------------------------------------------------------------------------
class ClockMonitor:
    """Samples SM clocks / throttle reasons DURING the timed region: NVML in a thread (a query takes well under a
    millisecond, so even a 50 ms region gets several samples); `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, enabled=True):
        # rank 0 only: eight processes polling NVML contend on the driver's lock and slow every rank's launches
        self.enabled = enabled
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []       # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()

    def _nvml_loop(self, handle):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(mhz), int(bits)))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if not self.enabled:
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map torch's (CUDA_VISIBLE_DEVICES-relative) index through the UUID
            uuid = torch.cuda.get_device_properties(self.index).uuid
            handle = None
            for cand in (f"GPU-{uuid}", str(uuid)):
                try:
                    handle = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                    break
                except Exception:
                    continue
            if handle is None:
                handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.nvml = nv
            # the launching thread holds the GIL almost continuously: let the sampler in every 0.5 ms
            self._switch = sys.getswitchinterval()
            sys.setswitchinterval(0.0005)
            self.thread = threading.Thread(target=self._nvml_loop, args=(handle,), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "rank 0 samples"}
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            sys.setswitchinterval(self._switch)
            nv = self.nvml
            masks = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [m for m, _ in self.samples]
            reasons = sorted(n for n, bit in masks.items() if any(b & bit for _, b in self.samples))
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(w, sample_units, steps, warmup, threads=None):
    """The reference's own op sequence on the host cores (oracle/torch_path.py): utt/s on a bounded
    sample of the workload (same T,U,V,J; fewer utterances)."""
    from oracle import torch_path

    if threads:
        torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(0)
    Bc = sample_units
    if w["kind"] == "rnnt":
        eouts = torch.randn(Bc, w["T"], w["He"], generator=gen, requires_grad=True)
        douts = torch.tanh(torch.randn(Bc, w["U"] + 1, w["Hd"], generator=gen)).requires_grad_()
        ys = make_labels(Bc, w["U"], w["V"], gen)
        tl = torch.full((Bc,), w["T"]); ul = torch.full((Bc,), w["U"])
        torch.manual_seed(1234)
        lin = [torch.nn.Linear(w["He"], w["J"]), torch.nn.Linear(w["Hd"], w["J"]), torch.nn.Linear(w["J"], w["V"])]

        def step():
            for l in lin:
                l.zero_grad(set_to_none=True)
            loss = torch_path.rnnt_joint_loss(eouts, douts, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias,
                                              lin[2].weight, lin[2].bias, ys, tl, ul, blank=0)
            loss.backward()
            return float(loss.detach())
    else:
        eouts = torch.randn(Bc, w["T"], w["He"], generator=gen, requires_grad=True)
        ys = make_labels(Bc, w["U"], w["V"], gen)
        tl = torch.full((Bc,), w["T"]); ul = torch.randint(w["U"] // 2, w["U"] + 1, (Bc,), generator=gen)
        torch.manual_seed(1234)
        head = torch.nn.Linear(w["He"], w["V"])

        def step():
            eouts.grad = None
            head.zero_grad(set_to_none=True)
            loss = torch_path.ctc_head_loss(eouts, head.weight, head.bias, ys, tl, ul, blank=0)
            loss.backward()
            return float(loss.detach())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return Bc * steps / dt, dt / steps


def emit(out):
    """The ONE line of stdout.  File descriptor 1 is pointed at stderr for the rest of the process (main()), so
    that banners printed by native libraries (e.g. "NCCL version ...") cannot add lines to it."""
    _REAL_STDOUT.write(json.dumps(out) + "\n")
    _REAL_STDOUT.flush()


_REAL_STDOUT = sys.stdout


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rnnt_cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--lengths", default="full", choices=["full", "ragged"])
    ap.add_argument("--grad-payload-mb", type=float, default=0.0,
                    help="N>1: MB of stand-in gradients (the rest of the model: 99.4 for the 25.8 M-parameter cfg-3 "
                         "model, 103 MB in total) all-reduced per step besides the path's own 3.7 MB.  They are "
                         "launched after the path's backward and nothing of the hot path is left to hide them "
                         "behind (in a training step they overlap the encoder backward), so this is the exposed "
                         "cost of the collective; default 0 = the path's own parameters only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--unfolded", action="store_true",
                    help="RNN-T workloads: w_enc / w_dec as cuBLAS Linear around the fused joint instead of inside the library")
    ap.add_argument("--ctc-unfused", action="store_true", help="CTC workloads: cuBLAS Linear + loss kernels on logits")
    ap.add_argument("--ctc-fused", action="store_true", help="CTC workloads: fused head also below 8192 frames")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "RNN-T/CTC loss fwd+bwd utterances/s"
    config = {"workload": w["desc"], "name": args.workload, "lengths": args.lengths,
              "per_gpu_batch": w["B"], "global_batch": w["B"] * world,
              "parallelism": (f"batch-sharded x{world}; NCCL all-reduce (AVG) of the path's parameter grads"
                              + (f" + {args.grad_payload_mb:.1f} MB stand-in for the rest of the 25.8 M-parameter model's grads"
                                 if args.grad_payload_mb > 0 and w["kind"] == "rnnt" else "")
                              + ", every gradient reduced in place from its autograd hook (overlaps the backward), stream-ordered wait")
              if world > 1 else "single GPU",
              "l2": "L2 flushed (256 MiB write, untimed) between timed steps",
              "e2e": "per step: H2D of the step's inputs from pinned host memory on a copy stream (prefetched "
                     "during the previous step) + D2H of the loss, read by the host one step late; wall clock; median of 3 "
                     "regions of `steps` steps each (all three under e2e.regions_utt_s_rank0), Python gc frozen first"}
    if w["kind"] == "rnnt":
        config["route"] = "logit tiles recomputed by the backward; dz through an L2-resident ring; no N x V tensor in HBM"
        config["projections"] = ("w_enc/w_dec (+ their backward) inside the library: tcgen05 GEMMs, bf16 operands"
                                 if args.precision == "bf16" and not args.unfolded else
                                 "w_enc/w_dec Linear via cuBLAS, " + ("TF32" if args.precision == "bf16" else "fp32"))
    else:
        config["head"] = ("output Linear(He,V) forward + backward included in the step: fused tensor-core head (no "
                          "(B,T,V) tensor in memory) where the shape allows, else cuBLAS TF32 + fp32 loss kernels")

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        # all host threads (torchrun exports OMP_NUM_THREADS=1, which would handicap the CPU arm)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else cores
        torch.set_num_threads(threads)
        config = dict(config, parallelism=f"host CPU, {threads} threads (rank 0 only)")
        config.pop("route", None)
        sample = 4 if w["kind"] == "rnnt" else w["B"]
        steps = max(1, min(args.steps, 3 if w["kind"] == "rnnt" else 5))
        warm = 1 if args.warmup > 0 else 0
        rate, sec = cpu_reference_rate(w, sample, steps, warm)
        out = {"impl": "reference", "metric": metric, "value": round(rate, 4), "unit": "utt/s", "n_gpus": args.gpus,
               "steps": steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": round(rate, 4), "unit": "utt/s", "cores": threads, "kind": "port",
                                "host_cpus": cores,
                                "sample": f"{sample} utterances/step of the same T,U,V,J shape, {steps} timed steps; "
                                          "reference op sequence (oracle/torch_path.py: joint->log_softmax->"
                                          "torchaudio rnnt_loss CPU / Linear->log_softmax->torch ctc_loss CPU)"},
               "e2e": {"value": round(rate, 4), "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return

    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    run = run_ours_rnnt if w["kind"] == "rnnt" else run_ours_ctc
    r = run(args, w, rank, world, dev)
    t = torch.tensor([r["ms_total"], r["e2e_s"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        units = r["units"] * world * args.steps
        value = units / (ms_total * 1e-3)
        out = {"metric": metric, "value": round(value, 2), "unit": "utt/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None,
               "dtype": args.precision if w["kind"] == "rnnt" else ("bf16" if (r["roofline"] or {}).get("head") == "fused" else "f32"),
               "data": "synthetic", "config": config, "clocks": r["clocks"],
               "e2e": {"value": round(units / e2e_s, 2), "unit": "utt/s", "h2d_bytes_per_step": r["h2d"],
                       "d2h_bytes_per_step": r["d2h"],
                       "regions_utt_s_rank0": [round(units / x, 1) for x in r["e2e_all"]]},
               "gpu_launches": r["launches"], "roofline": r["roofline"]}
        if r["flops"]:
            peaks = load_peaks()
            step_tf = r["flops"] / (ms_total / args.steps * 1e-3) / 1e12
            out["step_algorithmic_tflops"] = round(step_tf, 1)
            out["step_frac_of_burst_peak"] = round(step_tf / peaks["tf_burst"], 4)
            out["step_frac_of_sustained_peak"] = round(step_tf / (peaks["tf_sust"] or peaks["tf_burst"]), 4)
        extra = dict(r.get("extra") or {})
        if world == 1 and not args.no_extras:
            if args.workload == "rnnt_cfg3":
                # CTC cfg 2 through the same harness (its own roofline), so that it has a driver-run number too
                cargs = argparse.Namespace(**vars(args))
                cargs.workload, cargs.steps, cargs.warmup = "ctc_cfg2", max(5, args.steps // 2), 3
                c = run_ours_ctc(cargs, WORKLOADS["ctc_cfg2"], 0, 1, dev)
                cu = c["units"] * cargs.steps
                extra["ctc_cfg2"] = {"workload": WORKLOADS["ctc_cfg2"]["desc"],
                                     "value": round(cu / (c["ms_total"] * 1e-3), 1), "unit": "utt/s",
                                     "ms_per_step": round(c["ms_total"] / cargs.steps, 4),
                                     "e2e": {"value": round(cu / c["e2e_s"], 1), "unit": "utt/s",
                                             "h2d_bytes_per_step": c["h2d"], "d2h_bytes_per_step": c["d2h"],
                                             "regions_utt_s_rank0": [round(cu / x, 1) for x in c["e2e_all"]]},
                                     "roofline": c["roofline"], "dtype": "bf16" if c["roofline"].get("head") == "fused" else "f32"}
                torch.cuda.empty_cache()
            out["gpu_baseline"] = gpu_baselines(dev)
        if extra:
            out["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            # bounded sample of the same workload on the host cores: ~10-30 s of CPU work
            sample, csteps = (4, 2) if w["kind"] == "rnnt" else (min(w["B"], 16), 3)
            nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            rate, sec = cpu_reference_rate(w, sample, csteps, 1, threads=nthreads)
            out["cpu_baseline"] = {"value": round(rate, 4), "unit": "utt/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "host_cpus": os.cpu_count(),
                                   "sample": f"{sample} utterances of the same shape per step, 1 warm-up + {csteps} "
                                             f"timed steps ({sec:.1f} s/step); reference op sequence "
                                             "(oracle/torch_path.py)"}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

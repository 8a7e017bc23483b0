#!/usr/bin/env python
"""Benchmark of the sequence-loss hot path (BASELINE.json metric: RNN-T/CTC loss fwd+bwd utterances/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload rnnt_cfg3|rnnt_cfg4|ctc_cfg2|ctc_cfg1] [--precision bf16|fp32]
                    [--lengths full|ragged]

A "step" is one pass of the hot path over one batch of synthetic input:
  RNN-T  enc_proj = w_enc(eouts), dec_proj = w_dec(douts)  (plain cuBLAS Linear, as in the
         reference's joint) -> fused joint + log-softmax + transducer loss -> backward to eouts,
         douts and every joint parameter   (asr/modeling/decoders/rnn_transducer.py:101-115,147-156)
  CTC    logits -> fused log-softmax + CTC loss -> backward to logits (asr/modeling/decoders/ctc.py:109-113)
Default workload (N=1): BASELINE cfg 3, "RNN-T(Cf.) 1kBPE 26M fused joint+loss, B=32 T=250 U=100 V=1024".
For N>1 every rank processes its own batch of the same shape (weak scaling, batch-sharded) and the
gradients of the path's parameters are all-reduced over NCCL inside the timed step.

One JSON line is printed by rank 0 (see the driver contract in the task statement).
"""
import argparse
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu --set full captures under profiles/.
# recompute route (profiles/r1c_ncu_full_metrics.txt): bwd = joint_dh_kernel + reduce_dpre_kernel +
#   joint_dwt_kernel; algorithmic: h cache read twice, dpre written once and read once = 3.3 GB
# z-cache route (profiles/r1l_ncu_full_metrics.txt): fwd writes h + z; bwd = joint_dhz_kernel (z + w_out read,
#   dh written) + reduce_dh_tanh_kernel (dh read, h recomputed) + joint_dwz_kernel (z + h read); algorithmic:
#   z read twice, h read once, dh written and read once = 5.8 GB
NCU_DRAM_BYTES = {"rnnt_cfg3": {"fwd": 24.1e6 + 781.9e6,
                                "bwd": (842.4e6 + 776.5e6) + (834.2e6 + 21.9e6) + (844.1e6 + 4.5e6)},
                  "rnnt_cfg3_zc": {"fwd": 13.6e6 + 2451.2e6,
                                   "bwd": (1683.1e6 + 795.9e6) + (857.1e6 + 25.7e6) + (2498.1e6 + 3.9e6)}}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained"),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


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


def run_e2e(step, host, dev, steps, first=lambda out: out):
    """`steps` steps through the public API with HOST inputs: every step's inputs are copied from pinned host
    memory (on a copy stream, issued while the previous step computes -- what a prefetching loader does) and
    every step's loss is read back to the host.  Returns wall-clock seconds for exactly `steps` steps."""
    copy_stream = torch.cuda.Stream(device=dev)
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
    for i in range(steps):
        cur, ev = nxt
        main.wait_event(ev)
        for t in cur:
            t.record_stream(main)
        out = step(*cur)
        if i + 1 < steps:
            nxt = upload()
        float(first(out))
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t0


def run_ours_rnnt(args, w, rank, world, dev):
    import emoasr_b200 as E
    from emoasr_b200 import _lib, sharding

    # tensor-core mode: the two small projection GEMMs around the fused op (plain cuBLAS Linear, as in the
    # reference's joint) run in TF32 instead of SIMT fp32; fp32 mode keeps them in true fp32
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "bf16"
    wl = RNNTWorkload(w, seed=rank, regime=args.lengths)
    mods = torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
    params = list(mods.parameters())
    crit = E.RNNTJointLoss(blank_id=0, precision=args.precision)
    buckets = sharding.GradBuckets(params, own_grads=True) if world > 1 else None   # p.grad = views of flat buckets
    host = [t.pin_memory() for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(eouts, douts, ys, tlen, ulen):
        if buckets is not None:
            buckets.zero()          # in place: the gradients live inside the all-reduce buckets
        else:
            for p in params:
                p.grad = None
        eouts = eouts.detach().requires_grad_()
        douts = douts.detach().requires_grad_()
        loss = crit(wl.w_enc(eouts), wl.w_dec(douts), wl.output.weight, wl.output.bias, ys, tlen, ulen)
        loss.backward()
        if world > 1:
            buckets.allreduce()     # NCCL sum over ranks, / world (emoasr_b200/sharding.py)
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(*resident)
    barrier()
    # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between steps
    mon = ClockMonitor(dev.index if dev.index is not None else 0, enabled=rank == 0)
    mon.start()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(*resident)
        b.record()
        evs.append((a, b))
    barrier()
    clocks = mon.stop()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    # ---- end-to-end: host (pinned) buffers in, loss value out, wall clock
    run_e2e(step, host, dev, 2)
    barrier()
    e2e_s = run_e2e(step, host, dev, args.steps)
    barrier()

    # ---- dominant kernel alone: the fused joint forward (tcgen05) through the C ABI, CUDA events
    roof = None
    if args.precision == "bf16":
        import ctypes
        lib = _lib.load()
        with torch.no_grad():
            enc_proj = wl.w_enc(resident[0]).contiguous()
            dec_proj = wl.w_dec(resident[1]).contiguous()
        B, T, J = enc_proj.shape
        U1, V = dec_proj.size(1), w["V"]
        ws = torch.empty(_lib.workspace_bytes(0, 1, B, T, U1, J, V), dtype=torch.uint8, device=dev)
        lp2 = torch.empty(B, T, U1, 2, device=dev)
        lse = torch.empty(B, T, U1, device=dev)
        from emoasr_b200.functional import _joint_cache_bytes
        cache_bytes = _joint_cache_bytes(1, B, T, U1, J, V, dev)       # same policy as the autograd path
        zc = cache_bytes > _lib.workspace_bytes(_lib.OP_RNNT_JOINT_HCACHE, 1, B, T, U1, J, V)
        hc = torch.empty(cache_bytes, dtype=torch.uint8, device=dev)
        wo, bo = wl.output.weight.detach().contiguous(), wl.output.bias.detach().contiguous()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def call():
            rc = lib.emo_rnnt_joint_fwd(p(enc_proj), p(dec_proj), p(wo), p(bo), p(resident[2]), p(resident[3]),
                                        p(resident[4]), B, T, U1, J, V, 0, 1, p(lp2), p(lse), p(hc), hc.numel(),
                                        p(ws), ws.numel(), st)
            _lib.check(rc, "emo_rnnt_joint_fwd")
        # backward through the C ABI: lattice posteriors from the forward outputs, then emo_rnnt_joint_bwd
        alpha = torch.empty(B, T, U1, device=dev); beta = torch.empty(B, T, U1, device=dev)
        cost = torch.empty(B, device=dev); gamma2 = torch.empty(B, T, U1, 2, device=dev)
        call()
        _lib.check(lib.emo_rnnt_lattice_fwd_bwd(p(lp2), p(resident[3]), p(resident[4]), B, T, U1, p(alpha), p(beta),
                                                p(cost), p(gamma2), st), "emo_rnnt_lattice_fwd_bwd")
        gcost = torch.full((B,), 1.0 / B, device=dev)
        wsb = torch.empty(_lib.workspace_bytes(1, 1, B, T, U1, J, V), dtype=torch.uint8, device=dev)
        d_enc, d_dec = torch.empty_like(enc_proj), torch.empty_like(dec_proj)
        d_w, d_b = torch.empty_like(wo), torch.empty_like(bo)

        def call_bwd():
            rc = lib.emo_rnnt_joint_bwd(p(enc_proj), p(dec_proj), p(wo), p(bo), p(resident[2]), p(resident[3]),
                                        p(resident[4]), p(lse), p(gamma2), p(gcost), p(hc), hc.numel(), B, T, U1, J, V,
                                        0, 1, p(d_enc), p(d_dec), p(d_w), p(d_b), p(wsb), wsb.numel(), st)
            _lib.check(rc, "emo_rnnt_joint_bwd")

        def time_call(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            kev = []
            for _ in range(max(args.steps, 5)):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                kev.append((a, b))
            torch.cuda.synchronize()
            return statistics.mean(a.elapsed_time(b) for a, b in kev)

        f_ms, b_ms = time_call(call), time_call(call_bwd)
        peaks = load_peaks()
        unit = wl.joint_gemm_flops()

        def roofline(kernel, flops, ms, extra):
            ach = flops / (ms * 1e-3) / 1e12
            r = {"bound": "tensor", "kernel": kernel, "achieved": round(ach, 1), "peak": peaks["tf_burst"],
                 "unit": "TFLOP/s", "frac": round(ach / peaks["tf_burst"], 4), "traffic": None,
                 "kernel_ms": round(ms, 4), "peak_source": peaks["src"] + ", burst",
                 "algorithmic_flops_per_launch": flops}
            r.update(extra)
            return r
        # dominant by time: the backward call (ncu launch list under profiles/ gives the per-kernel shares)
        if zc:
            roof = roofline("emo_rnnt_joint_bwd = joint_dhz_kernel + joint_dwz_kernel (+ weight cast, axis reductions)",
                            2 * unit, b_ms, {"executed_flops_per_launch": 2 * unit,
                                             "note": "algorithmic = executed = dh and dW GEMMs (2 x 2*N*J*V); the "
                                                     "logits come from the fp16 z cache the forward wrote"})
        else:
            roof = roofline("emo_rnnt_joint_bwd = joint_dh_kernel + joint_dwt_kernel (+ weight cast, axis reductions)",
                            2 * unit, b_ms, {"executed_flops_per_launch": 6 * unit,
                                             "note": "algorithmic = dh and dW GEMMs (2 x 2*N*J*V); z is recomputed "
                                                     "per J-part in both kernels (6 GEMM units executed)"})
        roof["z_cache"] = bool(zc)
        roof["forward"] = roofline("joint_fwd_kernel (+ weight / stream casts)", unit, f_ms, {})
        # DRAM bytes per launch from the ncu --set full capture of this workload (profiles/, cfg 3 only)
        if args.workload == "rnnt_cfg3" and args.lengths == "full":
            key = "rnnt_cfg3_zc" if zc else "rnnt_cfg3"
            roof["traffic"] = NCU_DRAM_BYTES[key]["bwd"]
            roof["forward"]["traffic"] = NCU_DRAM_BYTES[key]["fwd"]

    prec = 1 if args.precision == "bf16" else 0
    per_step = (_lib.launch_count(_lib.OP_RNNT_JOINT_FWD, prec, w["B"], w["T"], w["U"] + 1, w["J"], w["V"]) +
                _lib.launch_count(_lib.OP_RNNT_JOINT_BWD, prec, w["B"], w["T"], w["U"] + 1, w["J"], w["V"]))
    launches = per_step * args.steps
    bytes_in = sum(t.numel() * t.element_size() for t in host)
    return dict(ms_total=ms_total, e2e_s=e2e_s, units=w["B"], clocks=clocks, roofline=roof, launches=launches,
                h2d=bytes_in, d2h=4, flops=wl.algorithmic_flops(), n_valid=wl.n_valid)


def run_ours_ctc(args, w, rank, world, dev):
    import emoasr_b200 as E
    from emoasr_b200 import _lib

    gen = torch.Generator().manual_seed(rank)
    B, T, U, V = w["B"], w["T"], w["U"], w["V"]
    logits_h = torch.randn(B, T, V, generator=gen).pin_memory()
    tlen, _ = make_lengths(B, T, U, args.lengths, gen)
    ulen = torch.randint(U // 2, U + 1, (B,), generator=gen)
    ys = make_labels(B, U, V, gen)
    host = [logits_h, ys.pin_memory(), tlen.pin_memory(), ulen.pin_memory()]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(logits, ys, tlen, ulen):
        logits = logits.detach().requires_grad_()
        loss = E.ctc_loss(logits, ys, tlen, ulen, blank=0, reduction="sum") / B
        loss.backward()
        return loss, logits.grad

    for _ in range(args.warmup):
        step(*resident)
    torch.cuda.synchronize()
    mon = ClockMonitor(dev.index if dev.index is not None else 0, enabled=rank == 0)
    mon.start()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(*resident); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = mon.stop()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    e2e_s = run_e2e(step, host, dev, args.steps, first=lambda out: out[0])
    peaks = load_peaks()
    alg_bytes = 3.0 * B * T * V * 4
    ach = alg_bytes / (ms_total / args.steps * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "ctc fwd+bwd (row_lse + lattice + grad kernels, whole step)",
            "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4),
            "traffic": None, "peak_source": peaks["src"], "algorithmic_bytes_per_step": alg_bytes}
    return dict(ms_total=ms_total, e2e_s=e2e_s, units=B, clocks=clocks, roofline=roof,
                launches=_lib.launch_count(_lib.OP_CTC, 0, B, T, 1, 1, V) * args.steps,
                h2d=sum(t.numel() * t.element_size() for t in host), d2h=4, flops=None, n_valid=None)


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
            return float(loss)
    else:
        logits = torch.randn(Bc, w["T"], w["V"], generator=gen, requires_grad=True)
        ys = make_labels(Bc, w["U"], w["V"], gen)
        tl = torch.full((Bc,), w["T"]); ul = torch.randint(w["U"] // 2, w["U"] + 1, (Bc,), generator=gen)

        def step():
            logits.grad = None
            loss = torch_path.ctc_loss_from_logits(logits, ys, tl, ul, blank=0)
            loss.backward()
            return float(loss)
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "RNN-T/CTC loss fwd+bwd utterances/s"
    config = {"workload": w["desc"], "name": args.workload, "lengths": args.lengths,
              "per_gpu_batch": w["B"], "global_batch": w["B"] * world,
              "parallelism": f"batch-sharded x{world}, NCCL all-reduce of the path's parameter grads" if world > 1 else "single GPU",
              "l2": "L2 flushed (256 MiB write, untimed) between timed steps",
              "projections": "w_enc/w_dec Linear via cuBLAS, " + ("TF32" if args.precision == "bf16" else "fp32"),
              "e2e": "per step: H2D of the step's inputs from pinned host memory on a copy stream (prefetched "
                     "during the previous step) + D2H read of the loss, wall clock"}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        # all host threads (torchrun exports OMP_NUM_THREADS=1, which would handicap the CPU arm)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else cores
        torch.set_num_threads(threads)
        config = dict(config, parallelism=f"host CPU, {threads} threads (rank 0 only)")
        sample = 2 if w["kind"] == "rnnt" else w["B"]
        steps = max(1, min(args.steps, 2 if w["kind"] == "rnnt" else 5))
        warm = 1 if args.warmup > 0 else 0
        rate, sec = cpu_reference_rate(w, sample, steps, warm)
        out = {"impl": "reference", "metric": metric, "value": round(rate, 4), "unit": "utt/s", "n_gpus": args.gpus,
               "steps": steps, "warmup": warm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": round(rate, 4), "unit": "utt/s", "cores": threads, "kind": "port",
                                "host_cpus": cores,
                                "sample": f"{sample} utterances/step of the same T,U,V,J shape, {steps} timed steps; "
                                          "reference op sequence (oracle/torch_path.py: joint->log_softmax->"
                                          "torchaudio rnnt_loss CPU / torch ctc_loss CPU)"},
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
               "scaling": "weak", "vs_baseline": None, "dtype": args.precision if w["kind"] == "rnnt" else "f32",
               "data": "synthetic", "config": config, "clocks": r["clocks"],
               "e2e": {"value": round(units / e2e_s, 2), "unit": "utt/s", "h2d_bytes_per_step": r["h2d"],
                       "d2h_bytes_per_step": r["d2h"]},
               "gpu_launches": r["launches"], "roofline": r["roofline"]}
        if r["flops"]:
            peaks = load_peaks()
            step_tf = r["flops"] / (ms_total / args.steps * 1e-3) / 1e12 * 1.0
            out["step_algorithmic_tflops"] = round(step_tf, 1)
            out["step_frac_of_sustained_peak"] = round(step_tf / (peaks["tf_sust"] or peaks["tf_burst"]), 4)
        if not args.no_cpu_baseline and world == 1:
            sample = 2 if w["kind"] == "rnnt" else min(w["B"], 16)
            nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            rate, sec = cpu_reference_rate(w, sample, 1, 1, threads=nthreads)
            out["cpu_baseline"] = {"value": round(rate, 4), "unit": "utt/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "host_cpus": os.cpu_count(),
                                   "sample": f"{sample} utterances of the same shape, 1 warm-up + 1 timed step "
                                             f"({sec:.1f} s/step)"}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

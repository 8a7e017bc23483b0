"""Where an end-to-end step's wall time goes: bench.run_e2e with per-step stamps.
For every step: host wall time of the iteration, device time of the H2D uploads (events on the copy stream) and of
the step's kernels (events on the compute stream).  usage: python tools/diag_e2e.py [--reps 4] [--steps 20] [--lengths full]"""
import argparse
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lengths", default="full")
    ap.add_argument("--workload", default="rnnt_cfg3")
    a = ap.parse_args()
    import emoasr_b200 as E
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w = bench.WORKLOADS[a.workload]
    wl = bench.RNNTWorkload(w, 0, a.lengths)
    torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
    params = [wl.output.weight, wl.output.bias, wl.w_dec.weight, wl.w_dec.bias, wl.w_enc.weight, wl.w_enc.bias]
    host = [t.pin_memory() for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
    print("pinned:", [t.is_pinned() for t in host], "bytes", sum(t.numel() * t.element_size() for t in host))
    resident = [t.to(dev) for t in host]
    step = bench.rnnt_step_fn(E, wl, "bf16", params, None, None, 1)
    for _ in range(5):
        step(*resident)
    torch.cuda.synchronize()
    # plain H2D bandwidth of the same buffers
    for _ in range(2):
        t0 = time.perf_counter()
        for _ in range(10):
            tens = [h.to(dev, non_blocking=True) for h in host]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print(f"plain H2D of the step's inputs: {dt * 1e3:.3f} ms  ({sum(t.numel() * t.element_size() for t in host) / dt / 1e9:.1f} GB/s)")
    for rep in range(a.reps):
        s = bench.run_e2e(step, host, dev, a.steps)
        print(f"rep {rep}: bench.run_e2e {s / a.steps * 1e3:.3f} ms/step")
    # instrumented copy of the loop
    for rep in range(a.reps):
        copy_stream = torch.cuda.Stream(device=dev)
        main_s = torch.cuda.current_stream(dev)
        pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        cev, sev, wall = [], [], []

        def upload():
            with torch.cuda.stream(copy_stream):
                e0 = torch.cuda.Event(enable_timing=True); e0.record(copy_stream)
                tens = [h.to(dev, non_blocking=True) for h in host]
                e1 = torch.cuda.Event(enable_timing=True); e1.record(copy_stream)
            cev.append((e0, e1))
            return tens, e1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nxt = upload()
        pending = None
        stamps = []
        for i in range(a.steps):
            ta = time.perf_counter()
            cur, ev = nxt
            main_s.wait_event(ev)
            for t in cur:
                t.record_stream(main_s)
            s0 = torch.cuda.Event(enable_timing=True); s0.record(main_s)
            out = step(*cur)
            s1 = torch.cuda.Event(enable_timing=True); s1.record(main_s)
            sev.append((s0, s1))
            tb = time.perf_counter()
            buf = pinned[i & 1]
            buf.copy_(out.detach().float(), non_blocking=True)
            done = torch.cuda.Event(); done.record(main_s)
            if i + 1 < a.steps:
                nxt = upload()
            tc = time.perf_counter()
            if pending is not None:
                pending[1].synchronize()
                float(pending[0])
            pending = (buf, done)
            td = time.perf_counter()
            stamps.append((tb - ta, tc - tb, td - tc))
        pending[1].synchronize()
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
        cms = [x.elapsed_time(y) for x, y in cev]
        sms = [x.elapsed_time(y) for x, y in sev]
        f = lambda v: f"min {min(v):.3f} med {statistics.median(v):.3f} max {max(v):.3f}"
        print(f"rep {rep}: {tot / a.steps * 1e3:.3f} ms/step | upload dev ms {f(cms)} | step dev ms {f(sms)} | "
              f"host launch ms {f([s[0] * 1e3 for s in stamps])} | host upload ms {f([s[1] * 1e3 for s in stamps])} | "
              f"host wait ms {f([s[2] * 1e3 for s in stamps])}")
    print("allocator:", torch.cuda.memory_stats(dev)["num_device_alloc"], "device allocs,",
          torch.cuda.memory_reserved(dev) / 1e6, "MB reserved")


if __name__ == "__main__":
    main()

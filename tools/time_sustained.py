"""Sustained (power-capped) step time of the default bench step: N back-to-back steps without the L2 flush, mean device
time of the last two thirds, NVML power / SM clock sampled meanwhile.  usage: python tools/time_sustained.py [steps]"""
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import emoasr_b200 as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = bench.WORKLOADS["rnnt_cfg3"]
wl = bench.RNNTWorkload(w, 0, "full")
torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
params = [wl.output.weight, wl.output.bias, wl.w_dec.weight, wl.w_dec.bias, wl.w_enc.weight, wl.w_enc.bias]
resident = [t.to(dev) for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
step = bench.rnnt_step_fn(E, wl, "bf16", params, None, None, 1)
for _ in range(5):
    step(*resident)
torch.cuda.synchronize()
samples, stop = [], threading.Event()


def poll():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        samples.append((nv.nvmlDeviceGetPowerUsage(h) / 1e3, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        time.sleep(0.02)


th = threading.Thread(target=poll, daemon=True)
th.start()
time.sleep(0.3)
evs = []
for _ in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(*resident); b.record()
    evs.append((a, b))
torch.cuda.synchronize()
stop.set(); th.join(timeout=1)
ms = [a.elapsed_time(b) for a, b in evs]
late = ms[n // 3:]
pw = [p for p, _ in samples[len(samples) // 3:]]
ck = [c for _, c in samples[len(samples) // 3:]]
print(f"sustained: first 10 steps {statistics.mean(ms[:10]):.3f} ms, last two thirds {statistics.mean(late):.3f} ms "
      f"(min {min(late):.3f} max {max(late):.3f}); power {statistics.mean(pw):.0f} W (max {max(pw):.0f}), "
      f"SM clock {statistics.mean(ck):.0f} MHz over {len(pw)} samples")

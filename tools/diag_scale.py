"""Per-step device times of the data-parallel step under torchrun: with the collective, without it, and the host
time of every iteration.  usage: torchrun --nproc-per-node N tools/diag_scale.py"""
import os, sys, time, statistics
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, emoasr_b200 as E  # noqa: E402
from emoasr_b200 import sharding  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOADS["rnnt_cfg3"]
wl = bench.RNNTWorkload(w, seed=rank, regime="full")
torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
params = [wl.output.weight, wl.output.bias, wl.w_dec.weight, wl.w_dec.bias, wl.w_enc.weight, wl.w_enc.bias]
resident = [t.to(dev) for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
red = sharding.GradReducer(params) if world > 1 else None
sync_step = bench.rnnt_step_fn(E, wl, "bf16", params, red, None, world)
local_step = bench.rnnt_step_fn(E, wl, "bf16", params, None, None, 1)


def run(step, n, do_flush=True):
    evs, host = [], []
    for _ in range(n):
        t0 = time.perf_counter()
        if do_flush:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(*resident); b.record()
        evs.append((a, b)); host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs], host


def show(tag, ms, host):
    print(f"rank {rank} {tag}: mean {statistics.mean(ms):.3f} median {statistics.median(ms):.3f} first {ms[0]:.3f} "
          f"min {min(ms):.3f} max {max(ms):.3f} | host/iter median {statistics.median(host):.3f} max {max(host):.3f}", flush=True)


import gc
for _ in range(5):
    sync_step(*resident)
torch.cuda.synchronize()
gc.collect(); gc.freeze()
for rep in range(2):
    ms, host = run(sync_step, 20); show("sync  20 flushed", ms, host)
    if red is not None:
        red.enabled = False
    ms, host = run(local_step, 20); show("local 20 flushed", ms, host)
    ms, host = run(local_step, 20, do_flush=False); show("local 20 no flush", ms, host)
    if red is not None:
        red.enabled = True
    ms, host = run(sync_step, 20, do_flush=False); show("sync  20 no flush", ms, host)
if world > 1:
    dist.destroy_process_group()

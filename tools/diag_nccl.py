"""Does an initialised NCCL process group change the device time of the (collective-free) local step?"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, emoasr_b200 as E  # noqa: E402

backend = sys.argv[1] if len(sys.argv) > 1 else "none"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
torch.backends.cuda.matmul.allow_tf32 = True
w = bench.WORKLOADS["rnnt_cfg3"]
wl = bench.RNNTWorkload(w, seed=rank, regime="full")
torch.nn.ModuleList([wl.w_enc, wl.w_dec, wl.output]).to(dev)
params = [wl.output.weight, wl.output.bias, wl.w_dec.weight, wl.w_dec.bias, wl.w_enc.weight, wl.w_enc.bias]
resident = [t.to(dev) for t in (wl.eouts, wl.douts, wl.ys.int(), wl.tlen.int(), wl.ulen.int())]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
step = bench.rnnt_step_fn(E, wl, "bf16", params, None, None, 1)
def measure(n=20):
    for _ in range(5): step(*resident)
    torch.cuda.synchronize()
    return bench.timed_steps(step, resident, n, flush, torch.cuda.synchronize) / n
before = measure()
if backend != "none":
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        t = torch.ones(1, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
after = measure()
print(f"rank {rank}/{world} backend {backend}: local step before init {before:.3f} ms, after {after:.3f} ms", flush=True)
if backend != "none":
    dist.barrier(); dist.destroy_process_group()

"""Times forward and backward of the fused CTC head at a BASELINE shape (CUDA events, L2 flushed)."""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emoasr_b200 as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--T", type=int, default=374)
ap.add_argument("--U", type=int, default=80)
ap.add_argument("--V", type=int, default=5000)
ap.add_argument("--He", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
eouts = torch.randn(a.B, a.T, a.He, generator=g).to(dev).requires_grad_()
lin = torch.nn.Linear(a.He, a.V).to(dev)
ys = torch.randint(4, a.V, (a.B, a.U), generator=g).to(dev)
tl = torch.full((a.B,), a.T, device=dev)
ul = torch.randint(a.U // 2, a.U + 1, (a.B,), generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
f_ms, b_ms = [], []
for it in range(a.iters + 3):
    lin.zero_grad(set_to_none=True)
    eouts.grad = None
    flush.fill_(1)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    loss = E.ctc_head_loss(eouts, lin.weight, lin.bias, ys, tl, ul, reduction="sum") / a.B
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    if it >= 3:
        f_ms.append(e0.elapsed_time(e1))
        b_ms.append(e1.elapsed_time(e2))
fm, bm = statistics.median(f_ms), statistics.median(b_ms)
print(f"head loss={float(loss):.4f} fwd {fm:.3f} ms  bwd {bm:.3f} ms  step {fm + bm:.3f} ms -> {a.B / (fm + bm) * 1e3:.0f} utt/s")

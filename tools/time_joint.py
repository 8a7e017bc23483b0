"""Times forward and backward of the fused joint (enc_proj / dec_proj in) at a BASELINE shape (CUDA events, L2 flushed)."""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emoasr_b200 as E  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--T", type=int, default=250)
ap.add_argument("--U", type=int, default=100)
ap.add_argument("--V", type=int, default=1024)
ap.add_argument("--J", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
enc = torch.randn(a.B, a.T, a.J, generator=g).to(dev).requires_grad_()
dec = torch.randn(a.B, a.U + 1, a.J, generator=g).to(dev).requires_grad_()
w = (torch.randn(a.V, a.J, generator=g) / a.J ** 0.5).to(dev).requires_grad_()
b = torch.zeros(a.V, device=dev, requires_grad=True)
ys = torch.randint(4, a.V, (a.B, a.U), generator=g).to(dev)
tl = torch.full((a.B,), a.T, device=dev)
ul = torch.full((a.B,), a.U, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
unit = 2.0 * a.B * a.T * (a.U + 1) * a.J * a.V
res = {}
for route in ("joint",):
    f_ms, b_ms = [], []
    for it in range(a.iters + 3):
        for t in (enc, dec, w, b):
            t.grad = None
        flush.fill_(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        loss = E.rnnt_joint_loss(enc, dec, w, b, ys, tl, ul, reduction="mean", precision="bf16")
        e1.record()
        loss.backward()
        e2.record()
        torch.cuda.synchronize()
        if it >= 3:
            f_ms.append(e0.elapsed_time(e1))
            b_ms.append(e1.elapsed_time(e2))
    fm, bm = statistics.median(f_ms), statistics.median(b_ms)
    res[route] = (float(loss), [t.grad.clone() for t in (enc, dec, w, b)])
    print(f"{route} loss={float(loss):.6f} fwd {fm:.3f} ms  bwd {bm:.3f} ms  step {fm + bm:.3f} ms "
          f"-> {a.B / (fm + bm) * 1e3:.0f} utt/s; algorithmic {3 * unit / (fm + bm) / 1e9:.0f} TFLOP/s")
if len(res) == 2:
    (l0, g0), (l1, g1) = res.values()
    print("loss diff", abs(l0 - l1), "grad rel diffs",
          [float((x - y).norm() / y.norm()) for x, y in zip(g0, g1)])

"""Runs the RNN-T hot path a few times at a BASELINE shape (for ncu / compute-sanitizer)."""
import argparse
import os
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
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--no-backward", action="store_true")
ap.add_argument("--ctc", action="store_true")
ap.add_argument("--ctc-head", action="store_true", help="fused CTC head from eouts (J = enc hidden size)")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if a.ctc_head:
    He = a.J
    eouts = torch.randn(a.B, a.T, He, generator=g).to(dev).requires_grad_()
    lin = torch.nn.Linear(He, a.V).to(dev)
    ys = torch.randint(4, a.V, (a.B, a.U), generator=g).to(dev)
    tl = torch.full((a.B,), a.T, device=dev)
    ul = torch.full((a.B,), a.U, device=dev)
    for _ in range(a.iters):
        lin.zero_grad(set_to_none=True)
        loss = E.ctc_head_loss(eouts, lin.weight, lin.bias, ys, tl, ul, reduction="sum") / a.B
        if not a.no_backward:
            loss.backward()
elif a.ctc:
    logits = torch.randn(a.B, a.T, a.V, generator=g).to(dev).requires_grad_()
    ys = torch.randint(4, a.V, (a.B, a.U), generator=g).to(dev)
    tl = torch.full((a.B,), a.T, device=dev)
    ul = torch.full((a.B,), a.U, device=dev)
    for _ in range(a.iters):
        logits.grad = None
        loss = E.ctc_loss(logits, ys, tl, ul, reduction="sum")
        if not a.no_backward:
            loss.backward()
else:
    enc = torch.randn(a.B, a.T, a.J, generator=g).to(dev).requires_grad_()
    dec = torch.randn(a.B, a.U + 1, a.J, generator=g).to(dev).requires_grad_()
    w = (torch.randn(a.V, a.J, generator=g) / a.J ** 0.5).to(dev).requires_grad_()
    b = torch.zeros(a.V, device=dev, requires_grad=True)
    ys = torch.randint(4, a.V, (a.B, a.U), generator=g).to(dev)
    tl = torch.full((a.B,), a.T, device=dev)
    ul = torch.full((a.B,), a.U, device=dev)
    for _ in range(a.iters):
        loss = E.rnnt_joint_loss(enc, dec, w, b, ys, tl, ul, reduction="mean", precision=a.precision)
        if not a.no_backward:
            loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))

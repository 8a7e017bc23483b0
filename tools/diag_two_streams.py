"""Two fused-joint training steps enqueued on two CUDA streams at the same time, repeated: the ring kernel needs all
of its CTAs resident at once, so two instances must never be interleaved on the SMs (cooperative launch).  Results are
compared with the same steps run one after the other.  Run under `timeout`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emoasr_b200 as E  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, T, U, V, J = 8, 250, 100, 1024, 512


def make(seed):
    g.manual_seed(seed)
    return [torch.randn(B, T, J, generator=g).to(dev), torch.randn(B, U + 1, J, generator=g).to(dev),
            (torch.randn(V, J, generator=g) / J ** 0.5).to(dev), torch.zeros(V, device=dev),
            torch.randint(1, V, (B, U), generator=g).to(dev)]


tl, ul = torch.full((B,), T, device=dev), torch.full((B,), U, device=dev)


def step(data):
    te = [t.clone().requires_grad_() for t in data[:4]]
    loss = E.rnnt_joint_loss(*te, data[4], tl, ul, blank=0, reduction="mean", precision="bf16")
    loss.backward()
    return [loss.detach()] + [t.grad for t in te]


a, b = make(1), make(2)
ref_a, ref_b = step(a), step(b)
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
bad = 0
for it in range(30):
    with torch.cuda.stream(s1):
        ra = step(a)
    with torch.cuda.stream(s2):
        rb = step(b)
    torch.cuda.synchronize()
    for got, ref in ((ra, ref_a), (rb, ref_b)):
        for x, y in zip(got, ref):
            if float((x - y).norm() / y.norm().clamp_min(1e-30)) > 1e-5:
                bad += 1
print("two-stream repeats done; mismatching tensors:", bad)

"""Run-to-run stress of the fused joint at a large ragged shape: every repeat is compared with the first.
    python tools/stress_repeat.py [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emoasr_b200 as E
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
routes = ["ring"]
B, T, U, V, J = 2, 1000, 400, 4096, 512
gen = torch.Generator().manual_seed(44)
enc = torch.randn(B, T, J, generator=gen).to(dev); dec = torch.randn(B, U + 1, J, generator=gen).to(dev)
w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev); bo = (0.1 * torch.randn(V, generator=gen)).to(dev)
ys = torch.randint(1, V, (B, U), generator=gen).to(dev)
tl, ul = torch.tensor([1000, 611], device=dev), torch.tensor([400, 333], device=dev)
ref, bad = {}, {r: 0 for r in routes}
for it in range(reps):
    for rt in routes:
        te = [t.clone().requires_grad_() for t in (enc, dec, w, bo)]
        costs = E.rnnt_joint_loss(*te, ys, tl, ul, blank=0, precision="bf16")
        costs.mean().backward(); torch.cuda.synchronize()
        g = [costs.detach().clone()] + [t.grad.clone() for t in te]
        if rt not in ref:
            ref[rt] = g
            continue
        errs = [float((a - b).norm() / b.norm()) for a, b in zip(g, ref[rt])]
        if max(errs) > 1e-4:
            bad[rt] += 1
            a, b = g[1], ref[rt][1]
            per = ((a - b).norm(dim=-1) / (b.norm(dim=-1) + 1e-12))
            print(it, rt, ["%.1e" % e for e in errs], "bad (b,t):", (per > 0.05).nonzero()[:6].tolist(), flush=True)
print("mismatching repeats:", bad, "of", reps - 1)

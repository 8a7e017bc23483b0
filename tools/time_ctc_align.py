"""CTC forced alignment at the cfg-2 shape (B=64, T=374, V=5000, U<=80): emo_ctc_align (one launch) against the
UNMODIFIED reference aligner (asr/modeling/decoders/ctc_aligner.py, staged under baseline/_ref/emoASR by
tools/stage_reference.py) on the same GPU tensors; checks that both give the same alignments."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import emoasr_b200 as E  # noqa: E402

B, T, V, U = 64, 374, 5000, 80
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
lp = torch.log_softmax(torch.randn(B, T, V, generator=g) * 2.0, dim=-1).to(dev)
ys = torch.randint(1, V, (B, U), generator=g).to(dev)
elens = torch.randint(T // 2, T + 1, (B,), generator=g)
elens[0] = T
ylens = torch.randint(U // 2, U + 1, (B,), generator=g)
el_d, yl_d = elens.to(dev), ylens.to(dev)
for _ in range(3):
    ours = E.ctc_forced_align(lp, ys, el_d, yl_d, blank=0)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    ours = E.ctc_forced_align(lp, ys, el_d, yl_d, blank=0)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
msg = f"ctc forced align B={B} T={T} V={V} U<={U}: emo_ctc_align {ms:.3f} ms"
ref_root = os.environ.get("EMOASR_REFERENCE") or os.path.join(ROOT, "baseline", "_ref", "emoASR")
if os.path.isdir(os.path.join(ref_root, "asr")):
    sys.path.insert(0, ref_root)
    from asr.modeling.decoders.ctc_aligner import CTCForcedAligner
    ref_al = CTCForcedAligner(blank_id=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref = ref_al(lp.clone(), elens, ys, ylens)
    torch.cuda.synchronize()
    rms = (time.perf_counter() - t0) * 1e3
    same = bool(torch.equal(ref.to(dev), ours))
    msg += f"; reference aligner on the same GPU {rms:.0f} ms; same alignments: {same}"
print(msg)

"""Greedy transducer search at the cfg-3 model size: batched on-device search (emo_rnnt_greedy_step) against the
reference's loop structure (per utterance, dense single-cell joint + argmax + .item() per step) on the same GPU."""
import os
import sys
import time
from collections import namedtuple

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emoasr_b200.decoders import RNNTDecoder  # noqa: E402

P = namedtuple("P", "dec_num_layers dec_hidden_size embedding_size joint_hidden_size enc_hidden_size vocab_size eos_id "
                    "blank_id mtl_ctc_weight kd_weight dropout_emb_rate dropout_dec_rate")
p = P(1, 512, 256, 512, 256, 1024, 2, 0, 0.0, 0, 0.0, 0.0)
B, T = 32, 250
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = RNNTDecoder(p, phase="test").to(dev).eval()
with torch.no_grad():
    dec.output.bias[0] += 3.0
eouts = torch.randn(B, T, p.enc_hidden_size, device=dev)
elens = torch.full((B,), T, device=dev)


def reference_style(n_utt):
    hyps = []
    with torch.no_grad():
        for b in range(n_utt):
            hyp = []
            ys = torch.full((1, 1), p.eos_id, dtype=torch.long, device=dev)
            dout, dstate = dec.recurrency(ys, None)
            t = 0
            while t < T:
                out = dec._dense_joint(eouts[b:b + 1, t:t + 1], dout)
                new_ys = out.squeeze(2).argmax(-1)
                tok = new_ys[0].item()
                if tok == p.blank_id:
                    t += 1
                else:
                    hyp.append(tok)
                    dout, dstate = dec.recurrency(new_ys, dstate)
                if len(hyp) > dec.max_seq_len:
                    break
            hyps.append(hyp)
    return hyps


dec._greedy(eouts, elens)
torch.cuda.synchronize()
t0 = time.perf_counter()
hyps, _, _, aligns = dec._greedy(eouts, elens)
torch.cuda.synchronize()
ours = time.perf_counter() - t0
n_ref = 4
reference_style(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
ref = reference_style(n_ref)
torch.cuda.synchronize()
theirs = (time.perf_counter() - t0) / n_ref * B
steps = max(len(a) for a in aligns)
print(f"greedy B={B} T={T}: batched on-device {ours * 1e3:.1f} ms ({steps} steps, {ours / steps * 1e6:.0f} us/step); "
      f"reference-style loop {theirs * 1e3:.0f} ms (extrapolated from {n_ref} utterances); same hyps: {hyps[:n_ref] == ref}")

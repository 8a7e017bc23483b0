"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference on CPU.

(ORACLE - test infrastructure.  Runs only in the build container: needs /root/reference.)

    python -m oracle.gen_golden            # writes tests/golden/ref_*.npz

What is executed is the reference's own code:
  * asr.modeling.decoders.rnn_transducer.RNNTDecoder.forward (rnn_transducer.py:81-145) with
    the warp_rnnt shim (oracle/warp_rnnt_shim.py; warp_rnnt itself is CUDA-only and absent),
  * asr.modeling.decoders.ctc.CTCDecoder.forward (ctc.py:87-174).
For each case the file stores the inputs, every parameter of the decoder (state_dict), the loss
values the reference returned and the gradients autograd produced for eouts and all parameters.
The two in-tree smoke inputs (rnnt_aligner.py:201-208, ctc_aligner.py:225-233) and the upstream
warp-transducer known-answer vector are frozen as fixtures as well.
"""
import os
import sys
from collections import namedtuple

import numpy as np
import torch

REF = os.environ.get("EMOASR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _params(**kw):
    base = dict(
        dec_num_layers=1, dec_hidden_size=16, embedding_size=8, joint_hidden_size=16,
        enc_hidden_size=12, vocab_size=11, eos_id=2, blank_id=0, mtl_ctc_weight=0.0,
        kd_weight=0, dropout_emb_rate=0.0, dropout_dec_rate=0.0,
    )
    base.update(kw)
    return namedtuple("Params", base.keys())(**base)


def _labels(g, B, U, V, ulens, eos):
    # labels drawn from {1} u [4,V) (specials per corpora/utils/spm_train.py:7-9), eos padded
    pool = torch.tensor([1] + list(range(4, V)))
    ys = pool[torch.randint(len(pool), (B, U), generator=g)]
    for b in range(B):
        ys[b, ulens[b]:] = eos
    return ys


def _np(t):
    return t.detach().cpu().numpy()


def rnnt_case(name, seed, B, T, U, p, tlens, ulens, mtl_ctc_weight=0.0):
    if not _wanted(name):
        return
    from asr.modeling.decoders.rnn_transducer import RNNTDecoder

    p = p._replace(mtl_ctc_weight=mtl_ctc_weight)
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    dec = RNNTDecoder(p, phase="test")
    dec.train()
    eouts = torch.randn(B, T, p.enc_hidden_size, generator=g).requires_grad_()
    elens = torch.tensor(tlens, dtype=torch.long)
    ylens = torch.tensor(ulens, dtype=torch.long)
    ys = _labels(g, B, U, p.vocab_size, ulens, p.eos_id)
    eos = torch.full((B, 1), p.eos_id, dtype=torch.long)
    ys_in = torch.cat([eos, ys], dim=1)          # datasets.py:160-173
    ys_out = torch.cat([ys, eos], dim=1)
    loss, loss_dict, logits = dec(eouts, elens, None, ys, ylens, ys_in, ys_out)
    loss.backward()
    out = {
        "eouts": _np(eouts), "elens": _np(elens), "ys": _np(ys), "ylens": _np(ylens),
        "ys_in": _np(ys_in), "ys_out": _np(ys_out),
        "loss_total": _np(loss), "grad_eouts": _np(eouts.grad), "logits": _np(logits),
        "meta_mtl_ctc_weight": np.float64(mtl_ctc_weight),
    }
    for k, v in loss_dict.items():
        out["lossdict." + k] = _np(v)
    for k, v in dec.state_dict().items():
        out["param." + k] = _np(v)
    for k, v in dec.named_parameters():
        out["grad." + k] = _np(v.grad) if v.grad is not None else np.zeros(0)
    for k, v in p._asdict().items():
        out["hp." + k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss))


def rnnt_kd_case(name, seed, B, T, U, p, tlens, ulens, kd_type, reduce_main):
    """RNNTDecoder.forward with knowledge distillation (rnn_transducer.py:127-141): kd_type "word"
    (RNNTWordDistillLoss, criteria.py:218-249) or "align" (RNNTForcedAligner + RNNTAlignDistillLoss,
    rnnt_aligner.py:155-198, criteria.py:252-288).  The aligner's Numba CUDA kernels run on the CPU under
    NUMBA_ENABLE_CUDASIM=1 (set by main())."""
    if not _wanted(name):
        return
    from asr.modeling.decoders.rnn_transducer import RNNTDecoder

    p = p._replace(kd_weight=0.3)
    p = namedtuple("Params", p._fields + ("kd_type", "reduce_main_loss_kd"))(*p, kd_type, reduce_main)
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    dec = RNNTDecoder(p, phase="train")
    dec.train()
    eouts = torch.randn(B, T, p.enc_hidden_size, generator=g).requires_grad_()
    elens = torch.tensor(tlens, dtype=torch.long)
    ylens = torch.tensor(ulens, dtype=torch.long)
    ys = _labels(g, B, U, p.vocab_size, ulens, p.eos_id)
    eos = torch.full((B, 1), p.eos_id, dtype=torch.long)
    ys_in = torch.cat([eos, ys], dim=1)
    ys_out = torch.cat([ys, eos], dim=1)
    soft_labels = torch.softmax(2.0 * torch.randn(B, U, p.vocab_size, generator=g), dim=-1)   # teacher posteriors
    loss, loss_dict, logits = dec(eouts, elens, None, ys, ylens, ys_in, ys_out, soft_labels)
    loss.backward()
    out = {
        "eouts": _np(eouts), "elens": _np(elens), "ys": _np(ys), "ylens": _np(ylens),
        "ys_in": _np(ys_in), "ys_out": _np(ys_out), "soft_labels": _np(soft_labels),
        "loss_total": _np(loss), "grad_eouts": _np(eouts.grad),
    }
    if kd_type == "align":
        with torch.no_grad():
            out["aligns"] = _np(dec.forced_aligner(torch.log_softmax(logits, dim=-1), elens, ys, ylens))
    for k, v in loss_dict.items():
        out["lossdict." + k] = _np(v)
    for k, v in dec.state_dict().items():
        out["param." + k] = _np(v)
    for k, v in dec.named_parameters():
        out["grad." + k] = _np(v.grad) if v.grad is not None else np.zeros(0)
    for k, v in p._asdict().items():
        out["hp." + k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss), {k: float(v) for k, v in loss_dict.items()})


def rnnt_greedy_case(name, seed, B, T, p, tlens, blank_bias):
    """Hypotheses and alignments of the reference's own greedy search (rnn_transducer.py:194-240, via decode()'s
    beam_width <= 1 branch) on a random model whose blank logit is biased so that blanks and labels alternate."""
    if not _wanted(name):
        return
    from asr.modeling.decoders.rnn_transducer import RNNTDecoder

    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    dec = RNNTDecoder(p, phase="test")
    dec.eval()
    with torch.no_grad():
        dec.output.bias[p.blank_id] += blank_bias
        dec.output.weight.mul_(3.0)             # peakier logits: fewer near-ties between fp32 summation orders
    eouts = torch.randn(B, T, p.enc_hidden_size, generator=g)
    elens = torch.tensor(tlens, dtype=torch.long)
    with torch.no_grad():
        hyps, scores, logits, aligns = dec._greedy(eouts, elens)
        # margin between the best and second-best logit at every step of the search (the test skips the comparison
        # of steps whose margin is below fp32 summation noise; with these weights there are none)
    out = {"eouts": _np(eouts), "elens": _np(elens), "n": np.asarray(B)}
    for b in range(B):
        out[f"hyp.{b}"] = np.asarray(hyps[b], dtype=np.int64)
        out[f"align.{b}"] = np.asarray(aligns[b], dtype=np.int64)
    for k, v in dec.state_dict().items():
        out["param." + k] = _np(v)
    for k, v in p._asdict().items():
        out["hp." + k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, [len(h) for h in hyps], [len(a) for a in aligns])


def ctc_case(name, seed, B, T, U, V, He, tlens, ulens):
    if not _wanted(name):
        return
    from asr.modeling.decoders.ctc import CTCDecoder

    p = _params(enc_hidden_size=He, vocab_size=V)
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    dec = CTCDecoder(p)
    eouts = torch.randn(B, T, He, generator=g).requires_grad_()
    elens = torch.tensor(tlens, dtype=torch.long)
    ylens = torch.tensor(ulens, dtype=torch.long)
    ys = _labels(g, B, U, V, ulens, p.eos_id)
    loss, loss_dict, logits = dec(eouts, elens, None, ys, ylens)
    loss.backward()
    out = {
        "eouts": _np(eouts), "elens": _np(elens), "ys": _np(ys), "ylens": _np(ylens),
        "loss_total": _np(loss), "grad_eouts": _np(eouts.grad), "logits": _np(logits),
    }
    for k, v in dec.state_dict().items():
        out["param." + k] = _np(v)
    for k, v in dec.named_parameters():
        out["grad." + k] = _np(v.grad)
    for k, v in p._asdict().items():
        out["hp." + k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss))


def ctc_mtl_case(name, seed, B, T, U, Up, V, Vp, He, tlens, ulens, plens, phone_w, hie, inter_w):
    if not _wanted(name):
        return
    """Phone-CTC (ctc.py:129-148, final or intermediate layer) and intermediate-CTC (ctc.py:150-170) heads."""
    from asr.modeling.decoders.ctc import CTCDecoder

    base = _params(enc_hidden_size=He, vocab_size=V)._asdict()
    base.update(mtl_phone_ctc_weight=phone_w, hie_mtl_phone=hie, phone_vocab_size=Vp, mtl_inter_ctc_weight=inter_w)
    p = namedtuple("Params", base.keys())(**base)
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    dec = CTCDecoder(p)
    eouts = torch.randn(B, T, He, generator=g).requires_grad_()
    eouts_inter = torch.randn(B, T, He, generator=g).requires_grad_()
    elens = torch.tensor(tlens, dtype=torch.long)
    ylens = torch.tensor(ulens, dtype=torch.long)
    pl = torch.tensor(plens, dtype=torch.long)
    ys = _labels(g, B, U, V, ulens, p.eos_id)
    ps = _labels(g, B, Up, Vp, plens, p.eos_id)
    loss, loss_dict, logits = dec(eouts, elens, eouts_inter, ys, ylens, None, None, None, ps, pl)
    loss.backward()
    out = {
        "eouts": _np(eouts), "eouts_inter": _np(eouts_inter), "elens": _np(elens), "ys": _np(ys), "ylens": _np(ylens),
        "ps": _np(ps), "plens": _np(pl), "loss_total": _np(loss), "grad_eouts": _np(eouts.grad),
        "grad_eouts_inter": _np(eouts_inter.grad) if eouts_inter.grad is not None else np.zeros(0),
        "logits": _np(logits),
    }
    for k, v in loss_dict.items():
        out["lossdict." + k] = _np(v)
    for k, v in dec.state_dict().items():
        out["param." + k] = _np(v)
    for k, v in dec.named_parameters():
        out["grad." + k] = _np(v.grad)
    for k, v in p._asdict().items():
        out["hp." + k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, float(loss), {k: float(v) for k, v in loss_dict.items()})


def smoke_fixtures():
    """Freeze the reference's two in-tree __main__ smoke inputs and what its CPU-runnable loss
    functions give on them, plus the warp-transducer known-answer vector."""
    import warp_rnnt  # the shim

    # rnnt_aligner.py:201-208 (seed 1, (2,10,6,5))
    torch.manual_seed(1)
    lp = torch.randn((2, 10, 6, 5)).log_softmax(dim=-1).requires_grad_()
    labels = torch.tensor([[1, 2, 1, 2, 0], [1, 2, 1, 2, 3]], dtype=torch.int32)
    T = torch.tensor([8, 10], dtype=torch.int32)
    U = torch.tensor([4, 5], dtype=torch.int32)
    costs = warp_rnnt.rnnt_loss(lp, labels, T, U, reduction=None, blank=0)
    costs.sum().backward()
    # the reference's forced aligner on the same input (Numba CUDA simulator on the CPU)
    from asr.modeling.decoders.rnnt_aligner import RNNTForcedAligner
    aligns = RNNTForcedAligner(blank_id=0)(lp.detach(), T, labels, U)
    np.savez_compressed(
        os.path.join(OUT, "ref_rnnt_aligner_smoke.npz"),
        log_probs=_np(lp), labels=_np(labels), T=_np(T), U=_np(U), costs=_np(costs), grad=_np(lp.grad),
        aligns=_np(aligns),
    )
    print("ref_rnnt_aligner_smoke", _np(costs))

    # ctc_aligner.py:225-233 (seed 1, (2,8,3))
    torch.manual_seed(1)
    logits = (torch.rand((2, 8, 3)) * 10.0).requires_grad_()
    elens = torch.tensor([7, 8])
    ys = torch.tensor([[1, 2, 0], [1, 2, 1]])
    ylens = torch.tensor([2, 3])
    fn = torch.nn.CTCLoss(blank=0, reduction="sum", zero_infinity=True)   # ctc.py:36-38
    loss = fn(logits.transpose(1, 0).log_softmax(dim=2), ys, elens, ylens) / logits.size(0)
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, "ref_ctc_aligner_smoke.npz"),
        logits=_np(logits), elens=_np(elens), ys=_np(ys), ylens=_np(ylens), loss=_np(loss),
        grad=_np(logits.grad),
    )
    print("ref_ctc_aligner_smoke", float(loss))

    # warp-transducer known answer: acts are *logits*
    acts = torch.tensor(
        [[[[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.6, 0.1, 0.1], [0.1, 0.1, 0.2, 0.8, 0.1]],
          [[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.2, 0.1, 0.1], [0.7, 0.1, 0.2, 0.1, 0.1]]]],
        requires_grad=True)
    lp = acts.log_softmax(-1)
    cost = warp_rnnt.rnnt_loss(lp, torch.tensor([[1, 2]], dtype=torch.int32),
                               torch.tensor([2], dtype=torch.int32), torch.tensor([2], dtype=torch.int32),
                               reduction=None, blank=0)
    cost.sum().backward()
    np.savez_compressed(
        os.path.join(OUT, "known_answer_warp_transducer.npz"),
        acts=_np(acts), labels=np.array([[1, 2]], np.int32), T=np.array([2], np.int32),
        U=np.array([2], np.int32), cost_published=np.float64(4.495666),
        grad_row_000_published=np.array([-0.13116688, -0.3999269, 0.17703125, 0.17703125, 0.17703125]),
        grad_row_012_published=np.array([-0.6925882, 0.16871116, 0.18645467, 0.16871116, 0.16871116]),
        cost_here=_np(cost), grad_here=_np(acts.grad),
    )
    print("known_answer", _np(cost))


def ctc_align_cases():
    """asr/modeling/decoders/ctc_aligner.py: the unmodified CTCForcedAligner on its own smoke input (:224-236) and on
    random ragged batches (repeated labels, empty label sequences, one infeasible utterance, peaked posteriors)."""
    from asr.modeling.decoders.ctc_aligner import CTCForcedAligner
    aligner = CTCForcedAligner(blank_id=0)
    out = {}

    def add(i, logits, elens, ys, ylens):
        lp = torch.log_softmax(logits, dim=-1)
        aligns = aligner(lp.clone(), elens, ys, ylens)      # the aligner zeroes padded frames of its argument
        out.update({f"c{i}_log_probs": _np(lp), f"c{i}_elens": _np(elens), f"c{i}_ys": _np(ys),
                    f"c{i}_ylens": _np(ylens), f"c{i}_aligns": _np(aligns)})

    torch.manual_seed(1)                                       # :225-233
    add(0, torch.rand((2, 8, 3)) * 10.0, torch.tensor([7, 8]), torch.tensor([[1, 2, 0], [1, 2, 1]]), torch.tensor([2, 3]))
    g = torch.Generator().manual_seed(21)
    add(1, torch.randn(4, 23, 11, generator=g) * 3.0, torch.tensor([23, 17, 9, 23]),
        torch.randint(1, 11, (4, 6), generator=g), torch.tensor([6, 4, 0, 2]))
    add(2, torch.randn(3, 40, 5, generator=g) * 6.0, torch.tensor([40, 31, 12]),       # 4 labels: many repeats
        torch.randint(1, 5, (3, 9), generator=g), torch.tensor([9, 9, 5]))
    add(3, torch.randn(3, 14, 29, generator=g), torch.tensor([14, 3, 14]),            # utterance 1: 3 frames, 5 labels
        torch.randint(1, 29, (3, 5), generator=g), torch.tensor([5, 5, 1]))
    add(4, torch.randn(5, 61, 203, generator=g) * 2.0, torch.tensor([61, 55, 50, 41, 30]),
        torch.randint(1, 203, (5, 20), generator=g), torch.tensor([20, 18, 20, 7, 11]))
    out["n_cases"] = np.int64(5)
    np.savez_compressed(os.path.join(OUT, "ref_ctc_forced_align.npz"), **out)
    print("ref_ctc_forced_align", [out[f"c{i}_aligns"].shape for i in range(5)])


ONLY = None   # python oracle/gen_golden.py NAME ... : regenerate just these cases


def _wanted(name):
    return ONLY is None or name in ONLY


def main():
    os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")   # rnnt_aligner.py's @cuda.jit kernels on the CPU
    global ONLY
    ONLY = set(sys.argv[1:]) or None
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import warp_rnnt_shim

    warp_rnnt_shim.install()
    sys.path.insert(0, REF)
    torch.set_num_threads(4)

    if ONLY is None or "smoke" in ONLY:
        smoke_fixtures()
    if ONLY is None or "ctc_align" in ONLY:
        ctc_align_cases()
    if ONLY is not None and ONLY <= {"smoke", "ctc_align"}:
        return
    p = _params()
    rnnt_case("ref_rnnt_small_full", 0, B=3, T=9, U=5, p=p, tlens=[9, 9, 9], ulens=[5, 5, 5])
    rnnt_case("ref_rnnt_small_ragged", 1, B=4, T=12, U=6, p=p, tlens=[12, 10, 7, 1], ulens=[6, 3, 0, 2])
    rnnt_case("ref_rnnt_small_auxctc", 2, B=3, T=14, U=4, p=p, tlens=[14, 11, 9], ulens=[4, 4, 2],
              mtl_ctc_weight=0.3)
    pm = _params(dec_num_layers=2, dec_hidden_size=40, embedding_size=24, joint_hidden_size=64,
                 enc_hidden_size=48, vocab_size=131)
    rnnt_case("ref_rnnt_medium_ragged", 3, B=5, T=37, U=17, p=pm,
              tlens=[37, 33, 30, 22, 15], ulens=[17, 17, 9, 12, 3])
    ctc_case("ref_ctc_small_full", 4, B=3, T=10, U=4, V=7, He=12, tlens=[10, 10, 10], ulens=[4, 4, 4])
    # includes an infeasible utterance (T_b < U_b: zero_infinity) and an empty label sequence
    ctc_case("ref_ctc_small_ragged", 5, B=5, T=13, U=6, V=9, He=12,
             tlens=[13, 9, 4, 13, 1], ulens=[6, 4, 6, 0, 1])
    ctc_case("ref_ctc_medium_ragged", 6, B=6, T=61, U=20, V=203, He=32,
             tlens=[61, 55, 50, 41, 30, 22], ulens=[20, 18, 20, 7, 11, 10])
    # tensor-core-compatible joint shape (J % 128 == 0, V % 32 == 0): the drop-in decoder in bf16 mode is
    # checked against the reference's own numbers at the stated bf16 tolerance
    pt = _params(dec_num_layers=1, dec_hidden_size=32, embedding_size=16, joint_hidden_size=128,
                 enc_hidden_size=24, vocab_size=64)
    rnnt_case("ref_rnnt_tcshape_ragged", 7, B=4, T=21, U=9, p=pt, tlens=[21, 18, 12, 5], ulens=[9, 6, 9, 1])
    rnnt_case("ref_rnnt_tcshape_auxctc", 8, B=3, T=17, U=6, p=pt, tlens=[17, 17, 10], ulens=[6, 2, 5],
              mtl_ctc_weight=0.3)
    # shape the FOLDED op takes (projections inside the library: He, Hd multiples of 16), with the auxiliary CTC head
    pf = _params(dec_num_layers=1, dec_hidden_size=32, embedding_size=16, joint_hidden_size=128,
                 enc_hidden_size=32, vocab_size=70)
    rnnt_case("ref_rnnt_tcfull_auxctc", 16, B=4, T=19, U=7, p=pf, tlens=[19, 19, 13, 6], ulens=[7, 4, 7, 2],
              mtl_ctc_weight=0.3)
    # phone-CTC on the final layer, on the intermediate layer (hie_mtl_phone), and intermediate CTC
    ctc_mtl_case("ref_ctc_phone_final", 9, B=4, T=19, U=5, Up=9, V=23, Vp=43, He=16, tlens=[19, 15, 12, 9],
                 ulens=[5, 4, 2, 5], plens=[9, 7, 3, 8], phone_w=0.3, hie=False, inter_w=0.0)
    ctc_mtl_case("ref_ctc_phone_hie_inter", 10, B=4, T=22, U=6, Up=10, V=29, Vp=43, He=16, tlens=[22, 20, 11, 6],
                 ulens=[6, 5, 6, 1], plens=[10, 8, 9, 2], phone_w=0.3, hie=True, inter_w=0.5)
    # knowledge distillation on the transducer (word / align), both ways of combining it with the main loss
    rnnt_kd_case("ref_rnnt_kd_word", 14, B=3, T=13, U=5, p=pm, tlens=[13, 11, 8], ulens=[5, 3, 4],
                 kd_type="word", reduce_main=True)
    rnnt_kd_case("ref_rnnt_kd_align", 15, B=3, T=12, U=5, p=pm, tlens=[12, 12, 7], ulens=[5, 4, 2],
                 kd_type="align", reduce_main=False)
    # the reference's greedy search (decode-time joint)
    rnnt_greedy_case("ref_rnnt_greedy", 13, B=5, T=23, p=pm, tlens=[23, 20, 14, 9, 1], blank_bias=3.0)
    # enc_hidden_size the fused tensor-core CTC head accepts (He % 128 == 0): main head with an odd vocabulary, an
    # infeasible utterance and an empty transcript; then all three heads (main, phone on the intermediate layer,
    # intermediate CTC)
    ctc_case("ref_ctc_tchead_ragged", 11, B=5, T=41, U=12, V=203, He=128,
             tlens=[41, 37, 9, 41, 20], ulens=[12, 10, 12, 0, 5])
    ctc_mtl_case("ref_ctc_tchead_phone_hie_inter", 12, B=4, T=33, U=8, Up=14, V=157, Vp=43, He=128,
                 tlens=[33, 30, 21, 16], ulens=[8, 7, 8, 2], plens=[14, 11, 12, 3], phone_w=0.3, hie=True, inter_w=0.5)


if __name__ == "__main__":
    main()

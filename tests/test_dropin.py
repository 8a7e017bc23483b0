"""Drop-in seams against the UNMODIFIED reference (CPU, build container only: needs /root/reference;
skipped elsewhere).  Checks what can be checked without a GPU: the three seams install, the patched
classes are subclasses of the reference's own decoders, state-dict keys / shapes are unchanged (so
checkpoints interchange), and the fused forward refuses to run on CPU instead of falling back."""
import os
import sys
from collections import namedtuple

import pytest
import torch

REF = os.environ.get("EMOASR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "asr")), reason="reference checkout not present")


def _params(**kw):
    base = dict(
        encoder_type="transformer", decoder_type="rnn_transducer", input_layer="conv2d", feat_dim=80, num_framestacks=1,
        enc_hidden_size=32, enc_num_attention_heads=2, enc_num_layers=1, enc_intermediate_size=64,
        dec_num_layers=1, dec_hidden_size=32, embedding_size=16, joint_hidden_size=128, vocab_size=64,
        eos_id=2, blank_id=0, mtl_ctc_weight=0.3, kd_weight=0, dropout_enc_rate=0.0, dropout_dec_rate=0.0,
        dropout_emb_rate=0.0, dropout_attn_rate=0.0, mtl_phone_ctc_weight=0, mtl_inter_ctc_weight=0,
    )
    base.update(kw)
    return namedtuple("Params", base.keys())(**base)


@pytest.fixture(scope="module")
def seams():
    from emoasr_b200 import dropin
    ref_asr = dropin.install(REF, precision="bf16")
    yield ref_asr
    for name in [m for m in sys.modules if m == "warp_rnnt"]:
        sys.modules.pop(name)


def test_module_seam_exposes_what_the_reference_touches(seams):
    import warp_rnnt                                    # rnn_transducer.py:14
    assert callable(warp_rnnt.rnnt_loss) and isinstance(warp_rnnt.__version__, str)   # :106-115, :65
    assert "emoasr_b200" in warp_rnnt.rnnt_loss.__module__


def test_class_seam_keeps_constructor_and_state_dict(seams):
    import asr.modeling.decoders.ctc as ref_ctc
    import asr.modeling.decoders.rnn_transducer as ref_rnnt
    p = _params()
    torch.manual_seed(0)
    fused = seams.RNNTDecoder(p, phase="train")         # looked up at call time by asr.py:40
    assert isinstance(fused, ref_rnnt.RNNTDecoder) and isinstance(fused.ctc, ref_ctc.CTCDecoder)
    assert type(fused).__mro__[1].__name__ == "FusedRNNTForward"
    import inspect
    plain_cls = [c for c in type(fused).__mro__ if c.__module__ == ref_rnnt.__name__][0]
    torch.manual_seed(0)
    plain = plain_cls(p, phase="train")
    sd_f, sd_p = fused.state_dict(), plain.state_dict()
    assert list(sd_f) == list(sd_p)
    assert all(sd_f[k].shape == sd_p[k].shape for k in sd_p)
    plain.load_state_dict(sd_f)                          # checkpoints interchange
    assert inspect.signature(type(fused).forward).parameters.keys() >= {
        "eouts", "elens", "eouts_inter", "ys", "ylens", "ys_in", "ys_out", "soft_labels", "ps", "plens"}
    # attribute seam: ctc_loss_fn is call-compatible with nn.CTCLoss
    from emoasr_b200.criteria import CTCLoss
    assert isinstance(fused.ctc.ctc_loss_fn, CTCLoss)


def test_fused_forward_has_no_cpu_fallback(seams):
    p = _params(mtl_ctc_weight=0.0)
    dec = seams.RNNTDecoder(p, phase="train")
    B, T, U = 2, 7, 3
    eouts = torch.randn(B, T, p.enc_hidden_size)
    ys = torch.randint(4, p.vocab_size, (B, U))
    ys_in = torch.cat([torch.full((B, 1), p.eos_id), ys], dim=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(eouts, torch.tensor([T, T - 1]), None, ys, torch.tensor([U, U]), ys_in, None)


def test_attribute_seam_replaces_the_ctc_forced_aligner(seams):
    """ctc.py:58-68: a distilling CTCDecoder builds a CTCForcedAligner; the seam swaps in the one-launch device version."""
    from emoasr_b200.criteria import CTCForcedAligner
    p = _params(decoder_type="ctc", kd_weight=0.5, lsm_prob=0.0, reduce_main_loss_kd=False)
    dec = seams.CTCDecoder(p)
    assert isinstance(dec.forced_aligner, CTCForcedAligner) and dec.forced_aligner.blank_id == p.blank_id
    assert not hasattr(seams.CTCDecoder(_params(decoder_type="ctc")), "forced_aligner")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec.forced_aligner(torch.zeros(1, 4, 5), torch.tensor([4]), torch.tensor([[1]]), torch.tensor([1]))


def test_ctc_greedy_decode_keeps_the_reference_results(seams):
    """ctc.py:176-200: the seam's _greedy (one host copy per batch) against the reference's per-frame .item() loop."""
    import asr.modeling.decoders.ctc as ref_ctc
    p = _params(decoder_type="ctc")
    torch.manual_seed(1)
    fused = seams.CTCDecoder(p)
    plain_cls = [c for c in type(fused).__mro__ if c.__module__ == ref_ctc.__name__][0]
    plain = plain_cls(p)
    plain.load_state_dict(fused.state_dict())
    eouts = torch.randn(3, 17, p.enc_hidden_size) * 3
    elens = torch.tensor([17, 11, 1])
    h0, s0, l0, a0 = plain._greedy(eouts, elens)
    h1, s1, l1, a1 = fused._greedy(eouts, elens)
    assert h0 == h1 and s0 == s1 and a0 == a1 and torch.equal(l0, l1)

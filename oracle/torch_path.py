"""The reference's own op sequence on CPU in torch fp32 (ORACLE / CPU baseline - test infrastructure).

This is the executable form of "the reference's CPU implementation of the path" that travels to
the GPU box (where /root/reference does not exist):

RNN-T  asr/modeling/decoders/rnn_transducer.py:147-156 (joint), :102 (log_softmax),
       :106-115 (warp_rnnt.rnnt_loss(..., reduction="mean", gather=False)).
       warp_rnnt is CUDA-only and absent from the image; torchaudio's CPU ``rnnt_loss`` with
       ``fused_log_softmax=False`` consumes log-probs and returns the same sparse gradient
       (checked against the fp64 DP in tests/test_oracle.py).
CTC    asr/modeling/decoders/ctc.py:103-115 (Linear -> transpose -> log_softmax ->
       nn.CTCLoss(blank, reduction="sum", zero_infinity=True) / B).
"""
import torch


def rnnt_loss_from_log_probs(log_probs, labels, frames_lengths, labels_lengths,
                             average_frames=False, reduction=None, blank=0, gather=False):
    """Signature of warp_rnnt.rnnt_loss as used at rnn_transducer.py:106-115."""
    import torchaudio.functional as AF

    assert not average_frames and not gather
    costs = AF.rnnt_loss(
        log_probs, labels.int(), frames_lengths.int(), labels_lengths.int(),
        blank=blank, reduction="none", fused_log_softmax=False,
    )
    if reduction == "mean":
        return costs.mean()
    if reduction == "sum":
        return costs.sum()
    return costs


def joint(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out):
    """rnn_transducer.py:147-156."""
    e = torch.nn.functional.linear(eouts.unsqueeze(2), w_enc, b_enc)
    d = torch.nn.functional.linear(douts.unsqueeze(1), w_dec, b_dec)
    out = torch.tanh(e + d)
    return torch.nn.functional.linear(out, w_out, b_out)


def rnnt_joint_loss(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out,
                    ys, elens, ylens, blank=0):
    """rnn_transducer.py:101-115: joint -> log_softmax -> rnnt_loss(mean)."""
    logits = joint(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out)
    log_probs = torch.log_softmax(logits, dim=-1)
    assert log_probs.size(2) == ys.size(1) + 1
    return rnnt_loss_from_log_probs(
        log_probs, ys.int(), elens.int(), ylens.int(),
        average_frames=False, reduction="mean", blank=blank, gather=False,
    )


def ctc_loss_from_logits(logits, ys, elens, ylens, blank=0):
    """ctc.py:109-113."""
    fn = torch.nn.CTCLoss(blank=blank, reduction="sum", zero_infinity=True)
    return fn(logits.transpose(1, 0).log_softmax(dim=2), ys, elens, ylens) / logits.size(0)


def ctc_head_loss(eouts, w, bias, ys, elens, ylens, blank=0):
    """ctc.py:103-113."""
    logits = torch.nn.functional.linear(eouts, w, bias)
    return ctc_loss_from_logits(logits, ys, elens, ylens, blank)

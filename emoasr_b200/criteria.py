"""Loss modules in the style of the reference's asr/criteria.py (nn.Module, ctor flags, padded
tensors + length tensors in, scalar out; template: asr/criteria.py:18-46)."""
import torch
from torch import nn

from . import functional as F


class RNNTLoss(nn.Module):
    """Transducer loss on dense log-probs -- what ``warp_rnnt.rnnt_loss(..., reduction="mean")``
    computes at asr/modeling/decoders/rnn_transducer.py:106-115."""

    def __init__(self, blank_id=0, normalize_length=False, normalize_batch=True):
        super().__init__()
        self.blank_id = blank_id
        self.normalize_length = normalize_length   # warp_rnnt's average_frames
        self.normalize_batch = normalize_batch     # reduction="mean" vs "sum"

    def forward(self, log_probs, ys, elens, ylens):
        return F.rnnt_loss(log_probs, ys, elens, ylens, average_frames=self.normalize_length,
                           reduction="mean" if self.normalize_batch else "sum", blank=self.blank_id)


class RNNTJointLoss(nn.Module):
    """Joint network + log-softmax + transducer loss in one fused op
    (rnn_transducer.py:101-115 and :147-156); the (B,T,U+1,V) tensors are never formed, forward or backward."""

    def __init__(self, blank_id=0, precision="bf16", normalize_length=False, normalize_batch=True):
        super().__init__()
        self.blank_id = blank_id
        self.precision = precision
        self.normalize_length = normalize_length
        self.normalize_batch = normalize_batch

    def forward(self, enc_proj, dec_proj, w_out, b_out, ys, elens, ylens):
        costs = F.rnnt_joint_loss(enc_proj, dec_proj, w_out, b_out, ys, elens, ylens,
                                  blank=self.blank_id, precision=self.precision)
        if self.normalize_length:
            costs = costs / elens.to(costs)
        return costs.mean() if self.normalize_batch else costs.sum()


class CTCLoss(nn.Module):
    """Call-compatible with the ``nn.CTCLoss`` instance the reference keeps in
    ``CTCDecoder.ctc_loss_fn`` (asr/modeling/decoders/ctc.py:36-38, called at :109-110 as
    ``fn(log_probs(T,B,V), ys, elens, ylens)``).  ``from_logits`` skips the transpose copy."""

    def __init__(self, blank=0, reduction="sum", zero_infinity=True):
        super().__init__()
        self.blank = blank
        self.reduction = reduction
        self.zero_infinity = zero_infinity

    def _finish(self, nll, target_lengths):
        if self.reduction == "sum":
            return nll.sum()
        if self.reduction == "mean":   # torch: divide by target length, then batch mean
            return (nll / target_lengths.to(nll).clamp_min(1)).mean()
        return nll

    def from_logits(self, logits, targets, input_lengths, target_lengths):
        """logits (B,T,V), raw or log-softmaxed."""
        nll = F.ctc_loss(logits, targets, input_lengths, target_lengths, blank=self.blank,
                         zero_infinity=self.zero_infinity)
        return self._finish(nll, target_lengths)

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        """log_probs (T,B,V) as nn.CTCLoss takes them."""
        return self.from_logits(log_probs.transpose(0, 1), targets, input_lengths, target_lengths)


class CTCHeadLoss(nn.Module):
    """Output Linear + log_softmax + CTC loss in one fused op (asr/modeling/decoders/ctc.py:103-113: ``logits =
    self.output(eouts)`` ... ``ctc_loss_fn(...) / B``); the (B,T,V) logits are never formed.  Takes the head's
    parameters at call time so that the module owning them (``CTCDecoder.output``) keeps its state-dict keys."""

    def __init__(self, blank=0, zero_infinity=True, normalize_batch=True):
        super().__init__()
        self.blank = blank
        self.zero_infinity = zero_infinity
        self.normalize_batch = normalize_batch   # the reference's "/ logits.size(0)" (ctc.py:111-113)

    def forward(self, eouts, weight, bias, ys, elens, ylens):
        nll = F.ctc_head_loss(eouts, weight, bias, ys, elens, ylens, blank=self.blank,
                              zero_infinity=self.zero_infinity)
        return nll.sum() / eouts.size(0) if self.normalize_batch else nll.sum()

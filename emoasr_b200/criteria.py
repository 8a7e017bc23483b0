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


class RNNTJointFullLoss(nn.Module):
    """RNNTJointLoss with the ``w_enc`` / ``w_dec`` projections (rnn_transducer.py:57-58,153), their backward and every
    cast inside the library: takes the encoder / prediction-network outputs and the three Linear layers' parameters.
    Tensor-core mode only."""

    def __init__(self, blank_id=0, normalize_batch=True):
        super().__init__()
        self.blank_id = blank_id
        self.normalize_batch = normalize_batch

    def forward(self, eouts, douts, w_enc, w_dec, output, ys, elens, ylens):
        costs = F.rnnt_joint_loss_from_outputs(eouts, douts, w_enc.weight, w_enc.bias, w_dec.weight, w_dec.bias,
                                               output.weight, output.bias, ys, elens, ylens, blank=self.blank_id)
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


class RNNTForcedAligner:
    """Call-compatible with asr/modeling/decoders/rnnt_aligner.py:155-198 (``aligner(log_probs, elens, ys, ylens)``
    -> best_aligns (B,U) int32): the Numba spin-lock alpha / beta kernels and the per-utterance Python walk become
    the wavefront lattice kernel and one on-device walk.  With the fused joint the alignment comes from the same
    lattice as the loss (``functional.rnnt_joint_outputs(..., aligns=True)``) and no dense tensor is needed."""

    def __init__(self, blank_id=0):
        self.blank_id = blank_id

    def __call__(self, log_probs, elens, ys, ylens):
        return F.rnnt_forced_align(log_probs.detach(), ys, elens, ylens, blank=self.blank_id)


class CTCForcedAligner:
    """Call-compatible with asr/modeling/decoders/ctc_aligner.py:88-221 (``aligner(log_probs, elens, ys, ylens)`` ->
    best_aligns (B,T) int64): the three Python loops over the frames (forward, backward, greedy pick with a per-frame
    argmax read back to the host) are one kernel launch for the batch.  ``dropin.install()`` puts it in place of every
    ``CTCDecoder.forced_aligner`` (the CTC distillation path, ctc.py:117-120,158-161)."""

    def __init__(self, blank_id=0):
        self.blank_id = blank_id

    def __call__(self, log_probs, elens, ys, ylens):
        return F.ctc_forced_align(log_probs, ys, elens, ylens, blank=self.blank_id)


class RNNTWordDistillLoss(nn.Module):
    """asr/criteria.py:218-249 without the dense (B,T,U+1,V) logits:
    ``-sum_{t<xlen,u<ylen} sum_v q[u,v] log_softmax(z[t,u])[v]`` with ``sum_v q log p = (q W).h + q.b - (sum q) lse``:
    ``lse`` is the fused joint's (differentiable) per-cell log-sum-exp, ``h = tanh(enc+dec)`` is formed one utterance
    at a time ((xlen, ylen, J): the reference keeps (B,T,U+1,J) AND three (B,T,U+1,V) tensors)."""

    def __init__(self, normalize_length=True, normalize_batch=True):
        super().__init__()
        self.normalize_length = normalize_length
        self.normalize_batch = normalize_batch

    def forward(self, enc_proj, dec_proj, w_out, b_out, lse, soft_labels, xlens, ylens):
        bs = enc_proj.size(0)
        loss = 0
        for b in range(bs):
            xlen, ylen = int(xlens[b]), int(ylens[b])
            q = soft_labels[b, :ylen].to(w_out.dtype)                    # (L, V)
            r = q @ w_out                                                # (L, J)
            h = torch.tanh(enc_proj[b, :xlen].unsqueeze(1) + dec_proj[b, :ylen].unsqueeze(0))   # (xlen, L, J)
            loss_b = (torch.einsum("tuj,uj->", h, r) + xlen * (q @ b_out).sum()
                      - (lse[b, :xlen, :ylen] * q.sum(-1).unsqueeze(0)).sum())
            if self.normalize_length:
                loss_b = loss_b / (xlen * ylen)
            loss = loss - loss_b
        if self.normalize_batch:
            loss = loss / bs
        return loss


class RNNTAlignDistillLoss(nn.Module):
    """asr/criteria.py:252-288 without the dense logits.  The reference's loop keeps only the LAST label's term
    (``loss_u`` is overwritten for u = 0 .. ylen-1 and subtracted once after the loop, :272-282); that is what its
    training runs optimise, so it is what this computes: one lattice cell (aligns[b, ylen-1], ylen-1) per utterance,
    whose logits are a (B,J) x (J,V) product."""

    def __init__(self, normalize_length=True, normalize_batch=True):
        super().__init__()
        self.normalize_length = normalize_length
        self.normalize_batch = normalize_batch

    def forward(self, enc_proj, dec_proj, w_out, b_out, soft_labels, aligns, xlens, ylens):
        bs = enc_proj.size(0)
        dev = enc_proj.device
        ylens = ylens.to(dev).long()
        if int(ylens.min()) < 1:
            raise RuntimeError("RNNTAlignDistillLoss: every utterance needs at least one label (as the reference)")
        bi = torch.arange(bs, device=dev)
        u = ylens - 1
        t = aligns.to(dev).long()[bi, u]
        h = torch.tanh(enc_proj[bi, t] + dec_proj[bi, u])                                 # (B, J)
        lp = torch.log_softmax(torch.nn.functional.linear(h, w_out, b_out), dim=-1)        # (B, V)
        loss_u = (soft_labels.to(dev)[bi, u].to(lp.dtype) * lp).sum(-1)
        if self.normalize_length:
            loss_u = loss_u / ylens.to(lp.dtype)
        loss = -loss_u.sum()
        if self.normalize_batch:
            loss = loss / bs
        return loss

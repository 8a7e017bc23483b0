"""Drop-in RNN-T / CTC decoders: same constructor, sub-module names (hence state_dict keys),
``forward`` signature and return values as asr/modeling/decoders/rnn_transducer.py:24-156 and
asr/modeling/decoders/ctc.py:26-174, with the loss path routed through the fused CUDA ops.

Two ways to use them:
  * stand-alone (``RNNTDecoder(params)``, ``CTCDecoder(params)``) -- training-time surface only
    (forward / joint / recurrency); this is what bench.py and the GPU tests drive.
  * ``emoasr_b200.dropin.install()`` mixes the ``Fused*Forward`` classes over the reference's own
    decoder classes, so decode / beam search / KD keep the reference's code.
"""
import torch
import torch.nn as nn

from . import functional as F
from .criteria import CTCLoss, RNNTAlignDistillLoss, RNNTWordDistillLoss


def _opt(params, name, default=0):
    return getattr(params, name) if hasattr(params, name) else default


def _reference_forward(self, mixin):
    """forward() of the first class behind `mixin` in the MRO that really implements one (the reference's
    decoder when installed through dropin.install()), or None for the stand-alone classes."""
    mro = type(self).__mro__
    for cls in mro[mro.index(mixin) + 1:]:
        fwd = cls.__dict__.get("forward")
        if fwd is not None and cls is not nn.Module:
            return fwd.__get__(self, type(self))
    return None


class FusedCTCForward:
    """forward() of ctc.py:87-174 with every ``ctc_loss_fn(logits.transpose(1,0).log_softmax(2), ...)
    / B`` site (:109-113, :139-141, :152-154) replaced by a fused op.

    ``fused_precision`` "fp32" (default of the stand-alone class): the Linear stays a cuBLAS call and the fused op
    takes its raw logits (log_softmax + loss, exact fp32).  "bf16" (what ``dropin.install()`` selects by default):
    Linear + log_softmax + loss + the Linear's backward run as ONE fused tensor-core op per head and the (B,T,V)
    logits are never formed -- the third return value is then None, like the RNN-T decoder's (its only reader is
    lm/modeling/p2w.py, outside this path); shapes the tensor-core path does not take fall back to the fp32 route."""

    fused_precision = "fp32"
    # below this many frames per batch the persistent tensor-core kernels cannot fill the GPU (a 256-frame tile per
    # CTA pair, 74 pairs) and the cuBLAS Linear + loss kernels are faster: cfg 1 (B=8, T=249) 0.33 vs 0.53 ms
    fused_head_min_frames = 8192

    def _fused_ctc(self, logits, ys, elens, ylens):
        nll = F.ctc_loss(logits, ys, elens, ylens, blank=self.blank_id, zero_infinity=True)
        return nll.sum() / logits.size(0)   # reduction="sum", then "/ B" (ctc.py:111-113)

    def _head_loss(self, linear, eouts, ys, elens, ylens):
        """(loss, logits or None) of one head (a Linear + CTC): fused from eouts when the tensor-core path takes it."""
        if (self.fused_precision == "bf16" and eouts.is_cuda and linear.bias is not None and
                eouts.size(0) * eouts.size(1) >= self.fused_head_min_frames and
                F.ctc_head_supported(eouts.size(0), eouts.size(1), eouts.size(2), linear.weight.size(0), ys.size(1))):
            nll = F.ctc_head_loss(eouts, linear.weight, linear.bias, ys, elens, ylens, blank=self.blank_id,
                                  zero_infinity=True)
            return nll.sum() / eouts.size(0), None
        logits = linear(eouts)
        return self._fused_ctc(logits, ys, elens, ylens), logits

    def forward(self, eouts, elens, eouts_inter=None, ys=None, ylens=None, ys_in=None, ys_out=None,
                soft_labels=None, ps=None, plens=None):
        wants_kd = self.kd_weight > 0 and soft_labels is not None
        if wants_kd or _opt(self, "inter_kd_weight") > 0:
            # KD needs dense log-probs + the forced aligner: the reference's own path
            ref_forward = _reference_forward(self, FusedCTCForward)
            if ref_forward is None:
                raise NotImplementedError("knowledge distillation needs the reference decoder "
                                          "(use emoasr_b200.dropin.install())")
            return ref_forward(eouts, elens, eouts_inter, ys, ylens, ys_in, ys_out, soft_labels, ps, plens)
        loss = 0
        loss_dict = {}
        if ys is None:
            return self.output(eouts)  # (B, T, vocab): decode-time use (ctc.py:105-106)
        loss_ctc, logits = self._head_loss(self.output, eouts, ys, elens, ylens)
        loss += loss_ctc
        loss_dict["loss_ctc"] = loss_ctc
        if self.mtl_phone_ctc_weight > 0:
            loss_phone_ctc, _ = self._head_loss(self.phone_output, eouts_inter if self.hie_mtl_phone else eouts,
                                                ps, elens, plens)
            loss += self.mtl_phone_ctc_weight * loss_phone_ctc
            key = "loss_phone_ctc(inter)" if self.hie_mtl_phone else "loss_phone_ctc"
            loss_dict[key] = loss_phone_ctc
        if self.mtl_inter_ctc_weight > 0:
            loss_inter_ctc, _ = self._head_loss(self.output, eouts_inter, ys, elens, ylens)
            loss_dict["loss_inter_ctc"] = loss_inter_ctc
            loss += self.mtl_inter_ctc_weight * loss_inter_ctc
        loss_dict["loss_total"] = loss
        return loss, loss_dict, logits


    def _greedy(self, eouts, elens, decode_phone=False):
        """ctc.py:176-200 with ONE device-to-host copy of the best path per batch instead of one ``.item()`` per
        frame and utterance (24 k host synchronisations at B=64, T=374); same return values."""
        from itertools import groupby
        logits = self.phone_output(eouts) if decode_phone else self.output(eouts)
        best_paths = logits.argmax(-1).cpu()
        lens = [int(n) for n in elens]
        hyps, scores, aligns = [], [], []
        for b in range(eouts.size(0)):
            indices = best_paths[b, :lens[b]].tolist()
            hyps.append([x for x, _ in groupby(indices) if x != self.blank_id])
            scores.append(None)
            aligns.append(indices)
        return hyps, scores, logits, aligns


class FusedRNNTForward:
    """forward() of rnn_transducer.py:81-145 with joint -> log_softmax -> warp_rnnt.rnnt_loss
    (:101-115) replaced by one fused op.  The third return value is None (the (B,T,U+1,V) logits are never
    formed; the only caller, asr/modeling/asr.py:65, discards it).

    ``fused_precision`` "bf16" runs the tensor-core kernels; shapes they do not support (see
    ``emo_rnnt_joint_supported``) fall back to the fp32 kernels with a one-time warning."""

    fused_precision = "bf16"
    _warned_fallback = False

    def _precision_for(self, B, T, U1, J, V):
        if self.fused_precision != "bf16" or F.joint_supported("bf16", B, T, U1, J, V):
            return self.fused_precision
        if not FusedRNNTForward._warned_fallback:
            FusedRNNTForward._warned_fallback = True
            import warnings
            warnings.warn(f"emoasr_b200: joint shape J={J}, V={V}, B={B} is outside what the bf16 tensor-core "
                          "kernels support; using the (much slower) fp32 kernels", RuntimeWarning)
        return "fp32"

    def forward(self, eouts, elens, eouts_inter=None, ys=None, ylens=None, ys_in=None, ys_out=None,
                soft_labels=None, ps=None, plens=None):
        loss = 0
        loss_dict = {}
        douts, _ = self.recurrency(ys_in, dstate=None)
        assert douts.size(1) == ys.size(1) + 1
        B, T, U1 = eouts.size(0), eouts.size(1), douts.size(1)
        J, V = self.output.weight.size(1), self.output.weight.size(0)
        precision = self._precision_for(B, T, U1, J, V)
        wants_kd = self.kd_weight > 0 and soft_labels is not None
        kd_type = getattr(self, "kd_type", None) if wants_kd else None
        if (precision == "bf16" and not wants_kd and eouts.is_cuda and self.w_enc.bias is not None and
                F.joint_full_supported(B, T, U1, eouts.size(2), douts.size(2), J, V)):
            # projections + joint + loss as two library calls (rnn_transducer.py:57-58,101-115,147-156)
            costs = F.rnnt_joint_loss_from_outputs(eouts, douts, self.w_enc.weight, self.w_enc.bias, self.w_dec.weight,
                                                   self.w_dec.bias, self.output.weight, self.output.bias, ys, elens,
                                                   ylens, blank=self.blank_id)
        else:
            enc_proj = self.w_enc(eouts)   # (B, T, J)
            dec_proj = self.w_dec(douts)   # (B, L+1, J)
            costs, lse, aligns = F.rnnt_joint_outputs(enc_proj, dec_proj, self.output.weight, self.output.bias, ys,
                                                      elens, ylens, blank=self.blank_id, precision=precision,
                                                      aligns=kd_type == "align")
        loss_rnnt = costs.mean()
        loss += loss_rnnt
        loss_dict["loss_rnnt"] = loss_rnnt
        if self.mtl_ctc_weight > 0:
            loss_ctc, _, _ = self.ctc(eouts=eouts, elens=elens, ys=ys, ylens=ylens, soft_labels=None)
            loss += self.mtl_ctc_weight * loss_ctc
            loss_dict["loss_ctc"] = loss_ctc
        if wants_kd:
            # rnn_transducer.py:127-141 on the fused lattice: the per-cell lse (word) / the forced alignment (align)
            # come out of the same fused forward as the loss; no (B,T,U+1,V) tensor is formed
            if kd_type == "word":
                loss_kd = RNNTWordDistillLoss()(enc_proj, dec_proj, self.output.weight, self.output.bias, lse,
                                                soft_labels, elens, ylens)
            elif kd_type == "align":
                loss_kd = RNNTAlignDistillLoss()(enc_proj, dec_proj, self.output.weight, self.output.bias, soft_labels,
                                                 aligns, elens, ylens)
            else:
                raise NotImplementedError(f"kd_type {kd_type!r} (the reference knows 'word' and 'align')")
            loss_dict["loss_kd"] = loss_kd
            if self.reduce_main_loss_kd:
                loss = (1 - self.kd_weight) * loss + self.kd_weight * loss_kd
            else:
                loss += self.kd_weight * loss_kd
        loss_dict["loss_total"] = loss
        return loss, loss_dict, None


class FusedRNNTSearch:
    """Decode-time side of the RNN-T decoder: the joint of the search loops
    (rnn_transducer.py:194-325) as one launch per step instead of (1,1,.) cuBLAS calls + tanh + argmax + .item().

    * ``joint()`` keeps the reference's signature; single-cell calls on CUDA without autograd -- what ``_greedy``
      and ``_beam_search`` make -- go through ``emo_rnnt_joint_step`` (the beam's hypotheses are the rows).
    * ``_greedy()`` is replaced by a batched search: all utterances advance together, the per-step joint + argmax +
      bookkeeping is ``emo_rnnt_greedy_step`` (no token is read back inside the loop), the prediction network runs
      batched for the rows that emitted.  Returns the reference's ``(hyps, scores, logits, aligns)``."""

    greedy_poll = 32   # steps between host polls of the "rows still active" counter

    def joint(self, eouts, douts):
        J = self.w_enc.weight.size(0)
        if (eouts.is_cuda and eouts.size(1) == 1 and douts.size(1) == 1 and J % 4 == 0 and
                not (torch.is_grad_enabled() and (eouts.requires_grad or douts.requires_grad or
                                                  self.output.weight.requires_grad)) and
                (eouts.size(0) == douts.size(0) or eouts.size(0) == 1 or douts.size(0) == 1)):
            with torch.no_grad():
                enc_proj = self.w_enc(eouts[:, 0])               # (Be, J)
                dec_proj = self.w_dec(douts[:, 0])               # (Bd, J)
                N = max(enc_proj.size(0), dec_proj.size(0))
                row = None
                if enc_proj.size(0) == 1 and N > 1:              # one frame against a beam of hypotheses
                    row = torch.zeros(N, dtype=torch.int32, device=eouts.device)
                if dec_proj.size(0) == 1 and N > 1:
                    dec_proj = dec_proj.expand(N, -1)
                logits, _ = F.joint_step(enc_proj, dec_proj, self.output.weight, self.output.bias, enc_row=row)
            return logits.view(N, 1, 1, -1)
        return self._dense_joint(eouts, douts)

    def _dense_joint(self, eouts, douts):
        mro = type(self).__mro__
        for cls in mro[mro.index(FusedRNNTSearch) + 1:]:       # the reference's own joint when installed over it
            j = cls.__dict__.get("joint")
            if j is not None:
                return j(self, eouts, douts)
        out = torch.tanh(self.w_enc(eouts.unsqueeze(2)) + self.w_dec(douts.unsqueeze(1)))   # rnn_transducer.py:150-154
        return self.output(out)

    @torch.no_grad()
    def _greedy(self, eouts, elens, decode_ctc_weight=0):
        if decode_ctc_weight == 1:
            return self.ctc.decode(eouts, elens, beam_width=1)
        B, T, _ = eouts.shape
        dev = eouts.device
        enc_proj = self.w_enc(eouts).contiguous()                                   # (B, T, J), once
        ys = torch.full((B, 1), self.eos_id, dtype=torch.long, device=dev)          # <sos>
        dout, dstate = self.recurrency(ys, None)
        state = F.GreedyState(B, T, torch.as_tensor(elens), self.max_seq_len, dev)
        w, b = self.output.weight.detach(), self.output.bias.detach()
        max_steps = T + self.max_seq_len + 2
        for step in range(1, max_steps + 1):
            dec_proj = self.w_dec(dout[:, 0]).contiguous()
            F.greedy_step(state, enc_proj, dec_proj, w, b, self.blank_id)
            # prediction network for the rows that emitted (run for all rows, kept where a token was appended)
            new_dout, new_dstate = self.recurrency(state.token.view(B, 1), dstate)
            em = state.emitted.bool()
            dout = torch.where(em.view(B, 1, 1), new_dout, dout)
            dstate = {k: torch.where(em.view(1, B, 1), new_dstate[k], dstate[k]) for k in ("hs", "cs")}
            if step % self.greedy_poll == 0 and int(state.n_active) == 0:
                break
        hyp, hyp_len = state.hyp.cpu(), state.hyp_len.cpu()
        align, align_len = state.align.cpu(), state.align_len.cpu()
        hyps = [hyp[i, :int(hyp_len[i])].tolist() for i in range(B)]
        aligns = [align[i, :int(align_len[i])].tolist() for i in range(B)]
        return hyps, [None] * B, None, aligns


class CTCDecoder(FusedCTCForward, nn.Module):
    """Stand-alone equivalent of asr/modeling/decoders/ctc.py:26-85 (constructor) for training."""

    def __init__(self, params):
        nn.Module.__init__(self)
        self.blank_id = params.blank_id
        self.eos_id = params.eos_id
        self.vocab_size = params.vocab_size
        self.output = nn.Linear(params.enc_hidden_size, self.vocab_size)
        self.ctc_loss_fn = CTCLoss(blank=self.blank_id, reduction="sum", zero_infinity=True)
        self.mtl_phone_ctc_weight = _opt(params, "mtl_phone_ctc_weight")
        self.mtl_inter_ctc_weight = _opt(params, "mtl_inter_ctc_weight")
        self.kd_weight = params.kd_weight
        if self.mtl_phone_ctc_weight > 0:
            self.hie_mtl_phone = params.hie_mtl_phone
            self.phone_output = nn.Linear(params.enc_hidden_size, params.phone_vocab_size)

    def decode(self, *args, **kwargs):
        raise NotImplementedError("decoding is outside the loss hot path: use the reference "
                                  "decoder via emoasr_b200.dropin.install()")


class RNNTDecoder(FusedRNNTForward, FusedRNNTSearch, nn.Module):
    """Stand-alone equivalent of asr/modeling/decoders/rnn_transducer.py:24-79 (constructor),
    :147-156 (joint) and :158-192 (recurrency) for training."""

    def __init__(self, params, phase="train"):
        nn.Module.__init__(self)
        self.dec_num_layers = params.dec_num_layers
        self.dec_hidden_size = params.dec_hidden_size
        self.eos_id = params.eos_id
        self.blank_id = params.blank_id
        self.max_seq_len = 256
        self.mtl_ctc_weight = params.mtl_ctc_weight
        self.kd_weight = params.kd_weight
        if self.kd_weight > 0 and phase == "train":          # rnn_transducer.py:66-79
            self.kd_type = params.kd_type
            self.reduce_main_loss_kd = params.reduce_main_loss_kd
        self.embed = nn.Embedding(params.vocab_size, params.embedding_size)
        self.dropout_emb = nn.Dropout(p=params.dropout_emb_rate)
        self.dropout = nn.Dropout(p=params.dropout_dec_rate)
        self.rnns = nn.ModuleList()
        input_size = params.embedding_size
        for _ in range(self.dec_num_layers):
            self.rnns += [nn.LSTM(input_size=input_size, hidden_size=params.dec_hidden_size,
                                  num_layers=1, batch_first=True)]
            input_size = params.dec_hidden_size
        self.w_enc = nn.Linear(params.enc_hidden_size, params.joint_hidden_size)
        self.w_dec = nn.Linear(params.dec_hidden_size, params.joint_hidden_size)
        self.output = nn.Linear(params.joint_hidden_size, params.vocab_size)
        if self.mtl_ctc_weight > 0:
            self.ctc = CTCDecoder(params)

    def recurrency(self, ys_in, dstate):
        ys_emb = self.dropout_emb(self.embed(ys_in))
        bs = ys_emb.size(0)
        if dstate is None:
            zeros = torch.zeros(self.dec_num_layers, bs, self.dec_hidden_size, device=ys_in.device)
            dstate = {"hs": zeros, "cs": zeros.clone()}
        new_hs, new_cs = [], []
        for layer_id in range(self.dec_num_layers):
            self.rnns[layer_id].flatten_parameters()
            ys_emb, (h, c) = self.rnns[layer_id](
                ys_emb, hx=(dstate["hs"][layer_id:layer_id + 1], dstate["cs"][layer_id:layer_id + 1]))
            new_hs.append(h)
            new_cs.append(c)
            ys_emb = self.dropout(ys_emb)
        return ys_emb, {"hs": torch.cat(new_hs, dim=0), "cs": torch.cat(new_cs, dim=0)}

    def decode(self, eouts, elens, eouts_inter=None, beam_width=1, len_weight=0, lm=None, lm_weight=0,
               decode_ctc_weight=0, decode_phone=False):
        """rnn_transducer.py:327-347: greedy search here (batched, on the device); beam search (ALSD with an optional
        LM) is the reference's Python -- install over it with emoasr_b200.dropin.install(), its per-step joint then
        runs through ``joint()`` above."""
        if beam_width > 1:
            raise NotImplementedError("beam search: use the reference decoder via emoasr_b200.dropin.install()")
        hyps, _, _, _ = self._greedy(eouts, elens, decode_ctc_weight)
        return hyps, None, None, None


"""Autograd functions over the C ABI (include/emoasr_b200.h).  CUDA tensors only; no fallback.

rnnt_loss        warp_rnnt.rnnt_loss signature (asr/modeling/decoders/rnn_transducer.py:106-115)
rnnt_joint_loss  joint + log_softmax + transducer loss fused (rnn_transducer.py:101-115, 147-156)
ctc_loss         log_softmax + nn.CTCLoss(reduction="none") fused (asr/modeling/decoders/ctc.py:109-113)
ctc_head_loss    output Linear + log_softmax + CTC loss fused, logits never formed (ctc.py:103-113)
"""
import ctypes

import math

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, 0: 0, 1: 1}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "emoasr_b200 ops run on CUDA tensors only (there is no CPU fallback); got a "
                f"{t.device} tensor")


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _i32c(t, device):
    return t.detach().to(device=device, dtype=torch.int32).contiguous()


def _i64c(t, device):
    return t.detach().to(device=device, dtype=torch.int64).contiguous()


# ----------------------------------------------------------------------------------------------
class _RNNTLattice(torch.autograd.Function):
    """cost(B) from gathered pairs lp2 (B,T,U1,2); d cost/d lp2 = -gamma2."""

    @staticmethod
    def forward(ctx, lp2, tlen, ulen):
        _require_cuda(lp2)
        lib = _lib.load()
        lp2c = _f32c(lp2)
        B, T, U1, two = lp2c.shape
        assert two == 2
        dev = lp2c.device
        tlen, ulen = _i32c(tlen, dev), _i32c(ulen, dev)
        with torch.cuda.device(dev):
            alpha = torch.empty(B, T, U1, device=dev)
            beta = torch.empty(B, T, U1, device=dev)
            cost = torch.empty(B, device=dev)
            gamma2 = torch.empty(B, T, U1, 2, device=dev)
            _lib.check(lib.emo_rnnt_lattice_fwd_bwd(_p(lp2c), _p(tlen), _p(ulen), B, T, U1, _p(alpha),
                                                    _p(beta), _p(cost), _p(gamma2), _stream()),
                       "emo_rnnt_lattice_fwd_bwd")
        ctx.save_for_backward(gamma2)
        return cost

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cost):
        (gamma2,) = ctx.saved_tensors
        return -gamma2 * grad_cost.view(-1, 1, 1, 1), None, None


class _RNNTDense(torch.autograd.Function):
    """warp_rnnt contract: dense log-probs in, sparse gradient out."""

    @staticmethod
    def forward(ctx, log_probs, labels, tlen, ulen, blank):
        _require_cuda(log_probs)
        lib = _lib.load()
        lp = _f32c(log_probs)
        B, T, U1, V = lp.shape
        dev = lp.device
        labels, tlen, ulen = _i32c(labels, dev), _i32c(tlen, dev), _i32c(ulen, dev)
        if labels.dim() != 2 or labels.size(0) != B or labels.size(1) != U1 - 1:
            raise RuntimeError(f"labels must be (B, U) = ({B}, {U1 - 1}); got {tuple(labels.shape)}")
        if U1 == 1:   # no labels at all: the kernels still want a valid pointer
            labels = torch.zeros(B, 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            lp2 = torch.empty(B, T, U1, 2, device=dev)
            alpha = torch.empty(B, T, U1, device=dev)
            beta = torch.empty(B, T, U1, device=dev)
            cost = torch.empty(B, device=dev)
            gamma2 = torch.empty(B, T, U1, 2, device=dev)
            _lib.check(lib.emo_rnnt_dense_fwd(_p(lp), _p(labels), _p(tlen), _p(ulen), B, T, U1, V, blank,
                                              _p(lp2), _p(alpha), _p(beta), _p(cost), _p(gamma2), _stream()),
                       "emo_rnnt_dense_fwd")
        ctx.save_for_backward(gamma2, labels, tlen, ulen)
        ctx.dims = (B, T, U1, V, blank)
        return cost

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cost):
        gamma2, labels, tlen, ulen = ctx.saved_tensors
        B, T, U1, V, blank = ctx.dims
        lib = _lib.load()
        dev = gamma2.device
        with torch.cuda.device(dev):
            g = _f32c(grad_cost)
            grad = torch.empty(B, T, U1, V, device=dev)
            _lib.check(lib.emo_rnnt_dense_bwd(_p(gamma2), _p(labels), _p(tlen), _p(ulen), _p(g), B, T, U1, V,
                                              blank, _p(grad), _stream()), "emo_rnnt_dense_bwd")
        return grad, None, None, None, None


def _reduce(costs, reduction):
    if reduction is None or reduction == "none":
        return costs
    if reduction == "mean":
        return costs.mean()
    if reduction == "sum":
        return costs.sum()
    raise ValueError(f"unknown reduction {reduction!r}")


def rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, average_frames=False,
              reduction=None, blank=0, gather=False):
    """Drop-in for ``warp_rnnt.rnnt_loss`` as called at rnn_transducer.py:106-115.

    log_probs (B,T,U+1,V) fp32 log-softmax output, labels (B,U) int, lengths (B,) int.  ``gather`` is
    warp_rnnt's memory switch (it gathers the {blank,label} pairs before the lattice): this implementation
    always gathers, so the flag changes nothing for dense input; as an extension, ``gather=True`` with a
    last dimension of 2 takes pre-gathered pairs (B,T,U+1,2).  Gradient w.r.t. log_probs has two
    non-zeros per valid lattice cell, exactly like warp_rnnt.
    """
    if gather and log_probs.size(-1) == 2:
        # extension: already gathered {blank, label} pairs (B,T,U+1,2)
        costs = _RNNTLattice.apply(log_probs, frames_lengths, labels_lengths)
    else:
        # dense log-probs: the kernel gathers the pairs itself (what warp_rnnt's gather=True does internally)
        costs = _RNNTDense.apply(log_probs, labels, frames_lengths, labels_lengths, int(blank))
    if average_frames:
        costs = costs / frames_lengths.to(costs)
    return _reduce(costs, reduction)


# ----------------------------------------------------------------------------------------------
def _padded_hidden(J):
    return (J + 127) // 128 * 128


def joint_supported(precision, B, T, U1, J, V):
    """True if rnnt_joint_loss can run these sizes with this precision (host call, no CUDA work).  The tensor-core
    kernels need J % 128 == 0: a joint size that is not is zero-padded to the next multiple by rnnt_joint_loss /
    rnnt_joint_outputs (see _pad_hidden), so what counts here is the padded size."""
    prec = _PRECISIONS[precision]
    return bool(_lib.load().emo_rnnt_joint_supported(prec, B, T, U1, _padded_hidden(J) if prec == _lib.PREC_BF16 else J, V))


def _pad_hidden(enc_proj, dec_proj, w_out):
    """Zero-pads the joint dimension to a multiple of 128 (tensor-core mode): tanh(0 + 0) = 0 meets zero columns of
    w_out, so logits, loss and the gradients of the real columns are unchanged; autograd slices the padding off."""
    pad = -enc_proj.size(-1) % 128
    if pad == 0:
        return enc_proj, dec_proj, w_out
    P = torch.nn.functional.pad
    return P(enc_proj, (0, pad)), P(dec_proj, (0, pad)), P(w_out, (0, pad))


class _RNNTJoint(torch.autograd.Function):
    """Outputs: cost (B); lse (B,T,U1), the log-sum-exp of every valid cell's logits (0 elsewhere) -- differentiable,
    so that losses built on log_softmax(z) without the dense tensor can be added (distillation: sum_v q log p =
    q.z - (sum q) lse); aligns (B,U) int32 forced alignment on the same lattice (only when asked for)."""

    @staticmethod
    def forward(ctx, enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, blank, precision, want_aligns):
        _require_cuda(enc_proj, dec_proj, w_out, b_out)
        lib = _lib.load()
        enc, dec, w, bo = _f32c(enc_proj), _f32c(dec_proj), _f32c(w_out), _f32c(b_out)
        B, T, J = enc.shape
        U1 = dec.size(1)
        V = w.size(0)
        if dec.size(0) != B or dec.size(2) != J or w.size(1) != J or bo.numel() != V:
            raise RuntimeError("rnnt_joint_loss: inconsistent shapes "
                               f"enc {tuple(enc.shape)} dec {tuple(dec.shape)} w_out {tuple(w.shape)}")
        dev = enc.device
        labels, tlen, ulen = _i32c(labels, dev), _i32c(tlen, dev), _i32c(ulen, dev)
        if U1 > 1:
            if labels.dim() != 2 or labels.size(0) != B or labels.size(1) < U1 - 1:
                raise RuntimeError(f"labels must be (B, >= U) = ({B}, {U1 - 1}); got {tuple(labels.shape)}")
            if labels.size(1) != U1 - 1:
                labels = labels[:, : U1 - 1].contiguous()
        else:
            labels = torch.zeros(B, 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            nbytes = _lib.workspace_bytes(_lib.OP_RNNT_JOINT_FWD, precision, B, T, U1, J, V)
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
            lp2 = torch.empty(B, T, U1, 2, device=dev)
            lse = torch.zeros(B, T, U1, device=dev)
            _lib.check(lib.emo_rnnt_joint_fwd(_p(enc), _p(dec), _p(w), _p(bo), _p(labels), _p(tlen), _p(ulen),
                                              B, T, U1, J, V, blank, precision, _p(lp2), _p(lse),
                                              _p(ws), ws.numel(), _stream()), "emo_rnnt_joint_fwd")
            alpha = torch.empty(B, T, U1, device=dev)
            beta = torch.empty(B, T, U1, device=dev)
            cost = torch.empty(B, device=dev)
            gamma2 = torch.empty(B, T, U1, 2, device=dev)
            _lib.check(lib.emo_rnnt_lattice_fwd_bwd(_p(lp2), _p(tlen), _p(ulen), B, T, U1, _p(alpha),
                                                    _p(beta), _p(cost), _p(gamma2), _stream()),
                       "emo_rnnt_lattice_fwd_bwd")
            aligns = torch.zeros(B, max(U1 - 1, 0), dtype=torch.int32, device=dev)
            if want_aligns and U1 > 1:
                _lib.check(lib.emo_rnnt_align(_p(alpha), _p(beta), _p(tlen), _p(ulen), B, T, U1, _p(aligns), _stream()),
                           "emo_rnnt_align")
        # nothing of size N x V is kept for the backward: it recomputes the logit tiles
        ctx.save_for_backward(enc, dec, w, bo, labels, tlen, ulen, lse, lp2, gamma2)
        ctx.cfg = (blank, precision)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(aligns)
        return cost, lse, aligns

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cost, grad_lse, _grad_aligns):
        enc, dec, w, bo, labels, tlen, ulen, lse, lp2, gamma2 = ctx.saved_tensors
        blank, precision = ctx.cfg
        lib = _lib.load()
        B, T, J = enc.shape
        U1, V = dec.size(1), w.size(0)
        dev = enc.device
        with torch.cuda.device(dev):
            g = _f32c(grad_cost) if grad_cost is not None else torch.zeros(B, device=dev)
            gl = None
            if grad_lse is not None:
                # only valid cells carry an lse: mask what autograd sends for the rest
                t_ok = torch.arange(T, device=dev).view(1, T, 1) < tlen.clamp(min=1).view(B, 1, 1)
                u_ok = torch.arange(U1, device=dev).view(1, 1, U1) <= ulen.clamp(min=0, max=U1 - 1).view(B, 1, 1)
                gl = (_f32c(grad_lse) * (t_ok & u_ok)).contiguous()
            nbytes = _lib.workspace_bytes(_lib.OP_RNNT_JOINT_BWD, precision, B, T, U1, J, V)
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
            d_enc = torch.empty_like(enc)
            d_dec = torch.empty_like(dec)
            d_w = torch.empty_like(w)
            d_b = torch.empty_like(bo)
            _lib.check(lib.emo_rnnt_joint_bwd(_p(enc), _p(dec), _p(w), _p(bo), _p(labels), _p(tlen), _p(ulen),
                                              _p(lse), _p(lp2), _p(gamma2), _p(g), _p(gl),
                                              B, T, U1, J, V, blank, precision,
                                              _p(d_enc), _p(d_dec), _p(d_w), _p(d_b), _p(ws), ws.numel(),
                                              _stream()), "emo_rnnt_joint_bwd")
        return d_enc, d_dec, d_w, d_b, None, None, None, None, None, None


def rnnt_joint_loss(enc_proj, dec_proj, w_out, b_out, labels, frames_lengths, labels_lengths,
                    blank=0, reduction=None, precision="bf16"):
    """Per-utterance transducer cost straight from the two projected streams.

    enc_proj (B,T,J) = w_enc(eouts)+b, dec_proj (B,U+1,J) = w_dec(douts)+b, w_out (V,J), b_out (V).
    Equivalent to ``rnnt_loss(log_softmax(output(tanh(enc_proj[:,:,None]+dec_proj[:,None]))), ...)``
    (rnn_transducer.py:101-115,147-156).  The (B,T,U+1,V) logits / log-probs / gradient are never formed,
    neither by the forward nor by the backward: precision="bf16" recomputes logit tiles on the tensor cores and
    hands ``dz`` to the gradient GEMMs through an L2-resident ring of tiles; precision="fp32" (parity mode)
    streams them through a bounded slab.
    """
    if _PRECISIONS[precision] == _lib.PREC_BF16:
        enc_proj, dec_proj, w_out = _pad_hidden(enc_proj, dec_proj, w_out)
    costs, _, _ = _RNNTJoint.apply(enc_proj, dec_proj, w_out, b_out, labels, frames_lengths,
                                   labels_lengths, int(blank), _PRECISIONS[precision], False)
    return _reduce(costs, reduction)


def rnnt_joint_outputs(enc_proj, dec_proj, w_out, b_out, labels, frames_lengths, labels_lengths, blank=0,
                       precision="bf16", aligns=False):
    """The fused op with all its outputs: ``(costs (B), lse (B,T,U+1), aligns (B,U) int32 or None)``.

    ``lse[b,t,u] = logsumexp_v z[b,t,u,v]`` for valid cells (0 elsewhere) is differentiable: together with a few
    per-cell dot products it gives any ``sum_v q[v] log_softmax(z)[v] = q.z - (sum q) lse`` without the dense
    tensor (knowledge distillation, asr/criteria.py:218-288).  ``aligns`` is the forced alignment of
    asr/modeling/decoders/rnnt_aligner.py:155-198 computed on the same lattice."""
    if _PRECISIONS[precision] == _lib.PREC_BF16:
        enc_proj, dec_proj, w_out = _pad_hidden(enc_proj, dec_proj, w_out)
    costs, lse, al = _RNNTJoint.apply(enc_proj, dec_proj, w_out, b_out, labels, frames_lengths,
                                      labels_lengths, int(blank), _PRECISIONS[precision], bool(aligns))
    return costs, lse, (al if aligns else None)


def joint_full_supported(B, T, U1, He, Hd, J, V):
    """True if rnnt_joint_loss_from_outputs can run these sizes (host call, no CUDA work)."""
    return bool(_lib.load().emo_rnnt_joint_full_supported(B, T, U1, He, Hd, J, V))


class _RNNTJointFull(torch.autograd.Function):
    """The fused joint with the w_enc / w_dec projections (and their backward) inside the library: two C calls per
    training step, working from the encoder / prediction-network outputs."""

    @staticmethod
    def forward(ctx, eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out, labels, tlen, ulen, blank):
        _require_cuda(eouts, douts, w_enc, w_dec, w_out)
        lib = _lib.load()
        e, d = _f32c(eouts), _f32c(douts)
        we, be, wd, bd, wo, bo = (_f32c(t) for t in (w_enc, b_enc, w_dec, b_dec, w_out, b_out))
        B, T, He = e.shape
        U1, Hd = d.size(1), d.size(2)
        J, V = we.size(0), wo.size(0)
        dev = e.device
        labels, tlen, ulen = _i32c(labels, dev), _i32c(tlen, dev), _i32c(ulen, dev)
        if U1 > 1:
            labels = labels[:, : U1 - 1].contiguous()
        else:
            labels = torch.zeros(B, 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            nbytes = int(lib.emo_rnnt_joint_full_workspace_bytes(0, B, T, U1, He, Hd, J, V))
            if nbytes == 0:
                raise RuntimeError(f"rnnt_joint_loss_from_outputs: unsupported shape B={B} T={T} U1={U1} He={He} Hd={Hd} "
                                   f"J={J} V={V} (see emo_rnnt_joint_full_supported)")
            fws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            lp2 = torch.empty(B, T, U1, 2, device=dev)
            lse = torch.empty(B, T, U1, device=dev)     # internal here: only valid cells are written and read back
            _lib.check(lib.emo_rnnt_joint_full_fwd(_p(e), _p(d), _p(we), _p(be), _p(wd), _p(bd), _p(wo), _p(bo), _p(labels),
                                                   _p(tlen), _p(ulen), B, T, U1, He, Hd, J, V, blank, _p(lp2), _p(lse),
                                                   _p(fws), fws.numel(), _stream()), "emo_rnnt_joint_full_fwd")
            alpha = torch.empty(B, T, U1, device=dev)
            beta = torch.empty(B, T, U1, device=dev)
            cost = torch.empty(B, device=dev)
            gamma2 = torch.empty(B, T, U1, 2, device=dev)
            _lib.check(lib.emo_rnnt_lattice_fwd_bwd(_p(lp2), _p(tlen), _p(ulen), B, T, U1, _p(alpha), _p(beta), _p(cost),
                                                    _p(gamma2), _stream()), "emo_rnnt_lattice_fwd_bwd")
        ctx.save_for_backward(bo, labels, tlen, ulen, lse, lp2, gamma2, fws)
        ctx.dims = (B, T, U1, He, Hd, J, V, blank)
        return cost

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_cost):
        bo, labels, tlen, ulen, lse, lp2, gamma2, fws = ctx.saved_tensors
        B, T, U1, He, Hd, J, V, blank = ctx.dims
        lib = _lib.load()
        dev = lse.device
        with torch.cuda.device(dev):
            g = _f32c(grad_cost)
            ws = torch.empty(int(lib.emo_rnnt_joint_full_workspace_bytes(1, B, T, U1, He, Hd, J, V)), dtype=torch.uint8,
                             device=dev)
            mk = lambda *s: torch.empty(*s, device=dev)
            d_e, d_d = mk(B, T, He), mk(B, U1, Hd)
            # the six parameter gradients are views of ONE flat buffer (16-byte aligned pieces): a data-parallel
            # reducer that sees them share a base all-reduces that buffer once (sharding.GradReducer)
            d_wo, d_bo, d_wd, d_bd, d_we, d_be = flat_views(dev, (V, J), (V,), (J, Hd), (J,), (J, He), (J,))
            _lib.check(lib.emo_rnnt_joint_full_bwd(_p(bo), _p(labels), _p(tlen), _p(ulen), _p(lse), _p(lp2), _p(gamma2),
                                                   _p(g), _p(None), _p(fws), B, T, U1, He, Hd, J, V, blank,
                                                   _p(d_e), _p(d_d), _p(d_we), _p(d_be), _p(d_wd), _p(d_bd), _p(d_wo),
                                                   _p(d_bo), _p(ws), ws.numel(), _stream()), "emo_rnnt_joint_full_bwd")
        return d_e, d_d, d_we, d_be, d_wd, d_bd, d_wo, d_bo, None, None, None, None


def flat_views(dev, *shapes):
    """fp32 tensors of the given shapes carved out of one flat allocation, every piece 16-byte aligned."""
    sizes = [(math.prod(sh) + 3) // 4 * 4 for sh in shapes]
    flat = torch.empty(sum(sizes), device=dev)
    out, off = [], 0
    for sh, n in zip(shapes, sizes):
        out.append(flat[off:off + math.prod(sh)].view(*sh))
        off += n
    return out


def rnnt_joint_loss_from_outputs(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out, labels, frames_lengths,
                                 labels_lengths, blank=0, reduction=None):
    """Per-utterance transducer cost from the encoder outputs (B,T,He) and the prediction-network outputs (B,U+1,Hd):
    ``rnnt_joint_loss(linear(eouts, w_enc, b_enc), linear(douts, w_dec, b_dec), w_out, b_out, ...)`` with the two
    projections, their weight / bias gradients and every cast inside the library (rnn_transducer.py:57-58,101-115,
    147-156).  Tensor-core mode (bf16 operands, fp32 accumulation); shapes outside ``joint_full_supported`` raise."""
    costs = _RNNTJointFull.apply(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out, labels, frames_lengths,
                                 labels_lengths, int(blank))
    return _reduce(costs, reduction)


def rnnt_forced_align(log_probs, labels, frames_lengths, labels_lengths, blank=0):
    """Drop-in for ``RNNTForcedAligner(blank_id)(log_probs, elens, ys, ylens)`` (rnnt_aligner.py:155-198) on dense
    log-probs (B,T,U+1,V): gathers the {blank,label} pairs, runs the lattice and walks it on the device.
    Returns best_aligns (B,U) int32."""
    _require_cuda(log_probs)
    lib = _lib.load()
    lp = _f32c(log_probs)
    B, T, U1, V = lp.shape
    dev = lp.device
    labels, tlen, ulen = _i32c(labels, dev), _i32c(frames_lengths, dev), _i32c(labels_lengths, dev)
    labels = labels[:, : U1 - 1].contiguous() if U1 > 1 else torch.zeros(B, 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        lp2 = torch.empty(B, T, U1, 2, device=dev)
        alpha, beta = torch.empty(B, T, U1, device=dev), torch.empty(B, T, U1, device=dev)
        cost, gamma2 = torch.empty(B, device=dev), torch.empty(B, T, U1, 2, device=dev)
        _lib.check(lib.emo_rnnt_dense_fwd(_p(lp), _p(labels), _p(tlen), _p(ulen), B, T, U1, V, int(blank), _p(lp2),
                                          _p(alpha), _p(beta), _p(cost), _p(gamma2), _stream()), "emo_rnnt_dense_fwd")
        aligns = torch.zeros(B, max(U1 - 1, 0), dtype=torch.int32, device=dev)
        _lib.check(lib.emo_rnnt_align(_p(alpha), _p(beta), _p(tlen), _p(ulen), B, T, U1, _p(aligns), _stream()),
                   "emo_rnnt_align")
    return aligns


# ----------------------------------------------------------------------------------------------
class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, tlen, ulen, blank, zero_infinity):
        _require_cuda(logits)
        lib = _lib.load()
        z = _f32c(logits)
        B, T, V = z.shape
        dev = z.device
        labels, tlen, ulen = _i64c(labels, dev), _i64c(tlen, dev), _i64c(ulen, dev)
        if labels.dim() != 2 or labels.size(0) != B:
            raise RuntimeError(f"labels must be (B, Umax); got {tuple(labels.shape)}")
        Umax = labels.size(1)
        if Umax == 0:
            labels = torch.zeros(B, 1, dtype=torch.int64, device=dev)
            Umax = 1
        S = 2 * Umax + 1
        with torch.cuda.device(dev):
            lse = torch.empty(B, T, device=dev)
            alpha = torch.empty(B, T, S, device=dev)
            # training step: the beta lattice runs side by side with alpha inside the forward call
            beta = torch.empty(B, T, S, device=dev) if ctx.needs_input_grad[0] else None
            nll = torch.empty(B, device=dev)
            _lib.check(lib.emo_ctc_fwd(_p(z), _p(labels), _p(tlen), _p(ulen), B, T, V, Umax, blank,
                                       int(zero_infinity), _p(lse), _p(alpha), _p(beta), _p(nll), _stream()),
                       "emo_ctc_fwd")
        ctx.save_for_backward(z, labels, tlen, ulen, lse, alpha, nll)
        ctx.beta = beta
        ctx.cfg = (blank, int(zero_infinity))
        return nll

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_nll):
        z, labels, tlen, ulen, lse, alpha, nll = ctx.saved_tensors
        blank, zero_infinity = ctx.cfg
        lib = _lib.load()
        B, T, V = z.shape
        Umax = labels.size(1)
        dev = z.device
        with torch.cuda.device(dev):
            g = _f32c(grad_nll)
            beta, beta_valid = ctx.beta, 1
            if beta is None:
                beta, beta_valid = torch.empty(B, T, 2 * Umax + 1, device=dev), 0
            grad = torch.empty_like(z)
            _lib.check(lib.emo_ctc_bwd(_p(z), _p(labels), _p(tlen), _p(ulen), _p(lse), _p(alpha), _p(nll),
                                       _p(g), B, T, V, Umax, blank, zero_infinity, _p(beta), beta_valid,
                                       _p(grad), _stream()), "emo_ctc_bwd")
        return grad, None, None, None, None, None


def ctc_loss(logits, labels, input_lengths, label_lengths, blank=0, reduction=None, zero_infinity=True):
    """Per-utterance CTC negative log-likelihood from RAW logits (B,T,V).

    Equivalent to ``nn.CTCLoss(blank, reduction="none", zero_infinity)(logits.transpose(1,0)
    .log_softmax(2), labels, input_lengths, label_lengths)`` (ctc.py:109-113); the log_softmax is
    fused.  Feeding log-probs instead of logits gives the same value and torch's gradient.
    """
    nll = _CTC.apply(logits, labels, input_lengths, label_lengths, int(blank), bool(zero_infinity))
    return _reduce(nll, reduction)


def ctc_forced_align(log_probs, labels, input_lengths, label_lengths, blank=0):
    """Drop-in for ``CTCForcedAligner(blank_id)(log_probs, elens, ys, ylens)`` (ctc_aligner.py:138-221): log_probs
    (B,T,V) = log_softmax(logits); returns best_aligns (B,T) int64 on the same device (the label or blank chosen for
    every frame, 0 past the utterance).  One launch; the argument is not modified (the reference zeroes its padded
    frames in place, which never influences the result)."""
    _require_cuda(log_probs)
    lib = _lib.load()
    lp = _f32c(log_probs.detach())
    B, T, V = lp.shape
    dev = lp.device
    labels, tlen, ulen = _i64c(labels, dev), _i64c(input_lengths, dev), _i64c(label_lengths, dev)
    if labels.dim() != 2 or labels.size(0) != B:
        raise RuntimeError(f"labels must be (B, Umax); got {tuple(labels.shape)}")
    if labels.size(1) == 0:
        labels = torch.zeros(B, 1, dtype=torch.int64, device=dev)
    Umax = labels.size(1)
    with torch.cuda.device(dev):
        aligns = torch.empty(B, T, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.emo_ctc_align_workspace_bytes(B, T, Umax)), dtype=torch.uint8, device=dev)
        _lib.check(lib.emo_ctc_align(_p(lp), _p(labels), _p(tlen), _p(ulen), B, T, V, Umax, int(blank), _p(aligns),
                                     _p(ws), ws.numel(), _stream()), "emo_ctc_align")
    return aligns


# ----------------------------------------------------------------------------------------------
def ctc_head_supported(B, T, He, V, Umax):
    """True if ctc_head_loss can run these sizes on the tensor-core path (host call, no CUDA work)."""
    return bool(_lib.load().emo_ctc_head_supported(B, T, _padded_hidden(He), V, max(Umax, 1)))   # see ctc_head_loss


class _CTCHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eouts, weight, bias, labels, tlen, ulen, blank, zero_infinity):
        _require_cuda(eouts, weight, bias)
        lib = _lib.load()
        e, w, bo = _f32c(eouts), _f32c(weight), _f32c(bias)
        B, T, He = e.shape
        V = w.size(0)
        if w.size(1) != He or bo.numel() != V:
            raise RuntimeError(f"ctc_head_loss: inconsistent shapes eouts {tuple(e.shape)} weight {tuple(w.shape)}")
        dev = e.device
        labels, tlen, ulen = _i64c(labels, dev), _i64c(tlen, dev), _i64c(ulen, dev)
        if labels.dim() != 2 or labels.size(0) != B:
            raise RuntimeError(f"labels must be (B, Umax); got {tuple(labels.shape)}")
        Umax = labels.size(1)
        if Umax == 0:
            labels = torch.zeros(B, 1, dtype=torch.int64, device=dev)
            Umax = 1
        S = 2 * Umax + 1
        need_grad = any(ctx.needs_input_grad[:3])
        with torch.cuda.device(dev):
            nbytes = int(lib.emo_ctc_head_workspace_bytes(0, B, T, He, V, Umax))
            if nbytes == 0:
                raise RuntimeError(f"ctc_head_loss: unsupported shape B={B} T={T} He={He} V={V} Umax={Umax} "
                                   "(see emo_ctc_head_supported; use ctc_loss on the Linear's output)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            lse = torch.zeros(B, T, device=dev)
            emis = torch.empty(B, T, Umax + 1, device=dev)
            alpha = torch.empty(B, T, S, device=dev)
            beta = torch.empty(B, T, S, device=dev) if need_grad else None
            nll = torch.empty(B, device=dev)
            _lib.check(lib.emo_ctc_head_fwd(_p(e), _p(w), _p(bo), _p(labels), _p(tlen), _p(ulen), B, T, He, V, Umax,
                                            blank, int(zero_infinity), _p(lse), _p(emis), _p(alpha), _p(beta), _p(nll),
                                            _p(ws), ws.numel(), _stream()), "emo_ctc_head_fwd")
        if need_grad:
            ctx.save_for_backward(e, w, bo, labels, tlen, ulen, lse, emis, alpha, beta)
        ctx.blank = blank
        return nll

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_nll):
        e, w, bo, labels, tlen, ulen, lse, emis, alpha, beta = ctx.saved_tensors
        lib = _lib.load()
        B, T, He = e.shape
        V, Umax = w.size(0), labels.size(1)
        dev = e.device
        with torch.cuda.device(dev):
            g = _f32c(grad_nll)
            ws = torch.empty(int(lib.emo_ctc_head_workspace_bytes(1, B, T, He, V, Umax)), dtype=torch.uint8, device=dev)
            d_e, d_w, d_b = torch.empty_like(e), torch.empty_like(w), torch.empty_like(bo)
            _lib.check(lib.emo_ctc_head_bwd(_p(e), _p(w), _p(bo), _p(labels), _p(tlen), _p(ulen), _p(lse), _p(emis),
                                            _p(alpha), _p(beta), _p(g), B, T, He, V, Umax, ctx.blank,
                                            _p(d_e), _p(d_w), _p(d_b), _p(ws), ws.numel(), _stream()),
                       "emo_ctc_head_bwd")
        return d_e, d_w, d_b, None, None, None, None, None


def ctc_head_loss(eouts, weight, bias, labels, input_lengths, label_lengths, blank=0, reduction=None,
                  zero_infinity=True):
    """Per-utterance CTC negative log-likelihood straight from the encoder outputs and the head's parameters.

    Equivalent to ``ctc_loss(linear(eouts, weight, bias), ...)`` (ctc.py:103-113) with the Linear, the log_softmax,
    the loss and all three backward GEMMs fused: the (B,T,V) logits are never formed.  Tensor-core path (bf16
    operands, fp32 accumulation); shapes outside ``ctc_head_supported`` raise.  ``He`` is zero-padded to a multiple of
    128 here when it is not one.
    """
    pad = -eouts.size(-1) % 128
    if pad:   # an encoder width that is not a multiple of 128: zero columns on both operands leave the logits unchanged
        eouts = torch.nn.functional.pad(eouts, (0, pad))
        weight = torch.nn.functional.pad(weight, (0, pad))
    nll = _CTCHead.apply(eouts, weight, bias, labels, input_lengths, label_lengths, int(blank), bool(zero_infinity))
    return _reduce(nll, reduction)


# ----------------------------------------------------------------------------------------------
def step_workspace(n_rows, device):
    """Zeroed workspace of the decode-step calls for up to n_rows rows (they leave it zeroed: allocate once per search)."""
    return torch.zeros(int(_lib.load().emo_rnnt_step_workspace_bytes(int(n_rows))), dtype=torch.uint8, device=device)


def joint_step(enc_proj, dec_proj, w_out, b_out, enc_row=None, want_logits=True, want_token=False, ws=None):
    """Decode-time joint for N rows in one launch (rnn_transducer.py:147-156 at T = L = 1, fp32):
    ``z[n] = w_out tanh(enc_proj[enc_row[n]] + dec_proj[n]) + b_out``.  enc_proj (rows,J), dec_proj (N,J), enc_row (N)
    int32 or None (row n).  Returns (logits (N,V) or None, token (N) int64 = argmax or None).  No grad."""
    _require_cuda(enc_proj, dec_proj, w_out, b_out)
    lib = _lib.load()
    enc, dec, w, bo = _f32c(enc_proj), _f32c(dec_proj), _f32c(w_out), _f32c(b_out)
    N, J = dec.shape
    V = w.size(0)
    dev = dec.device
    with torch.cuda.device(dev):
        logits = torch.empty(N, V, device=dev) if want_logits else None
        token = torch.empty(N, dtype=torch.int64, device=dev) if want_token else None
        if want_token and ws is None:
            ws = step_workspace(N, dev)
        row = _i32c(enc_row, dev) if enc_row is not None else None
        _lib.check(lib.emo_rnnt_joint_step(_p(enc), _p(row), _p(dec), _p(w), _p(bo), N, J, V, _p(logits), _p(token),
                                           _p(ws), ws.numel() if ws is not None else 0, _stream()), "emo_rnnt_joint_step")
    return logits, token


class GreedyState:
    """Device-side state of a batched greedy search (one row per utterance); see emo_rnnt_greedy_step."""

    def __init__(self, B, T, tlen, max_len, device, keep_align=True):
        self.B, self.T, self.max_len = B, T, max_len
        self.tlen = tlen.detach().to(device=device, dtype=torch.int32).clamp(min=0, max=T).contiguous()
        z = lambda *s, dt=torch.int32: torch.zeros(*s, dtype=dt, device=device)
        self.t_idx, self.hyp, self.hyp_len = z(B), z(B, max_len + 1), z(B)
        self.align_cap = T + max_len + 2 if keep_align else 0
        self.align = z(B, self.align_cap) if keep_align else None
        self.align_len = z(B)
        self.emitted = z(B, dt=torch.uint8)
        self.token = z(B, dt=torch.int64)
        self.n_active = z(1)
        self.ws = step_workspace(B, device)


def greedy_step(state, enc_proj, dec_proj, w_out, b_out, blank):
    """One step of batched greedy transducer search for all utterances (rnn_transducer.py:194-240): joint at every
    row's current frame, argmax, and the bookkeeping (advance on blank, append otherwise), on the device."""
    lib = _lib.load()
    B, T, J = enc_proj.shape
    V = w_out.size(0)
    s = state
    with torch.cuda.device(enc_proj.device):
        _lib.check(lib.emo_rnnt_greedy_step(_p(enc_proj), _p(dec_proj), _p(w_out), _p(b_out), _p(s.tlen), B, T, J, V,
                                            int(blank), s.max_len, _p(s.t_idx), _p(s.hyp), _p(s.hyp_len), _p(s.align),
                                            _p(s.align_len), s.align_cap, _p(s.emitted), _p(s.token), _p(s.n_active),
                                            _p(s.ws), s.ws.numel(), _stream()), "emo_rnnt_greedy_step")

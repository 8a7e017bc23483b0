"""fp64 numpy restatement of the RNN-T joint + transducer loss (ORACLE - test infrastructure).

Follows the reference's arithmetic, not its code:

* joint      asr/modeling/decoders/rnn_transducer.py:147-156
             z[b,t,u,:] = W_out . tanh(W_enc e[b,t] + b_enc + W_dec d[b,u] + b_dec) + b_out
* log-probs  asr/modeling/decoders/rnn_transducer.py:102   lp = log_softmax(z, -1)
* loss       asr/modeling/decoders/rnn_transducer.py:106-115
             warp_rnnt.rnnt_loss(lp, ys, elens, ylens, average_frames=False,
                                 reduction="mean", blank, gather=False)
             (third party: 1ytic/warp-rnnt, unpinned; published algorithm = Graves 2012
             forward-backward, same recursion the reference spells out in
             asr/modeling/decoders/rnnt_aligner.py:49-83 (alpha) and :121-152 (beta)).

Everything is float64 and written as explicit dynamic programming over anti-diagonals.
"""
import numpy as np

NEG_INF = -np.inf


def _logaddexp(a, b):
    return np.logaddexp(a, b)


def lattice(lp_blank, lp_label, T, U):
    """Alpha/beta over one utterance.

    lp_blank[t,u] = log P(blank | t,u), lp_label[t,u] = log P(y_{u+1} | t,u)  (u < U)
    shapes (Tmax, U1max); only [:T, :U+1] is read.
    Returns alpha, beta (T, U+1) and ll = log P(y|x).
    rnnt_aligner.py:49-83 / :121-152 (without its /T[b] normalisation).
    """
    U1 = U + 1
    alpha = np.full((T, U1), NEG_INF)
    beta = np.full((T, U1), NEG_INF)
    alpha[0, 0] = 0.0
    for d in range(1, T + U1 - 1):
        u_lo = max(0, d - (T - 1))
        u_hi = min(U, d)
        u = np.arange(u_lo, u_hi + 1)
        t = d - u
        no_emit = np.full(u.shape, NEG_INF)
        emit = np.full(u.shape, NEG_INF)
        m = t > 0
        no_emit[m] = alpha[t[m] - 1, u[m]] + lp_blank[t[m] - 1, u[m]]
        m = u > 0
        emit[m] = alpha[t[m], u[m] - 1] + lp_label[t[m], u[m] - 1]
        alpha[t, u] = _logaddexp(no_emit, emit)
    ll = alpha[T - 1, U] + lp_blank[T - 1, U]

    beta[T - 1, U] = lp_blank[T - 1, U]
    for d in range(T + U1 - 3, -1, -1):
        u_lo = max(0, d - (T - 1))
        u_hi = min(U, d)
        u = np.arange(u_lo, u_hi + 1)
        t = d - u
        no_emit = np.full(u.shape, NEG_INF)
        emit = np.full(u.shape, NEG_INF)
        m = t < T - 1
        no_emit[m] = beta[t[m] + 1, u[m]] + lp_blank[t[m], u[m]]
        m = u < U
        emit[m] = beta[t[m], u[m] + 1] + lp_label[t[m], u[m]]
        beta[t, u] = _logaddexp(no_emit, emit)
    return alpha, beta, ll


def occupancies(lp_blank, lp_label, T, U):
    """cost = -ll and the two per-cell transition posteriors.

    gamma_blank[t,u] = exp(alpha[t,u] + lp_blank[t,u] + beta[t+1,u] - ll)  (beta[T,U] := 0)
    gamma_label[t,u] = exp(alpha[t,u] + lp_label[t,u] + beta[t,u+1] - ll)  (0 at u == U)
    d cost / d lp_blank = -gamma_blank,  d cost / d lp_label = -gamma_label
    (the sparse gradient contract of warp_rnnt).
    Arrays are (Tmax, U1max), zero outside the valid region.
    """
    alpha, beta, ll = lattice(lp_blank, lp_label, T, U)
    gb = np.zeros_like(lp_blank, dtype=np.float64)
    gl = np.zeros_like(lp_blank, dtype=np.float64)
    if not np.isfinite(ll):
        return -ll, gb, gl, alpha, beta
    U1 = U + 1
    beta_next_t = np.full((T, U1), NEG_INF)
    beta_next_t[:-1] = beta[1:]
    beta_next_t[T - 1, U] = 0.0
    gb[:T, :U1] = np.exp(alpha + lp_blank[:T, :U1] + beta_next_t - ll)
    if U > 0:
        gl[:T, :U] = np.exp(alpha[:, :U] + lp_label[:T, :U] + beta[:, 1:] - ll)
    return -ll, gb, gl, alpha, beta


def rnnt_loss_dense(log_probs, labels, tlens, ulens, blank=0):
    """warp_rnnt.rnnt_loss(..., reduction=None, gather=False) restated.

    log_probs (B,T,U1,V); labels (B,U) ; returns costs (B,) and d costs/d log_probs (sparse,
    two non-zeros per valid cell).
    """
    log_probs = np.asarray(log_probs, dtype=np.float64)
    B, Tm, U1m, V = log_probs.shape
    costs = np.zeros(B)
    grad = np.zeros_like(log_probs)
    for b in range(B):
        T, U = int(tlens[b]), int(ulens[b])
        lpb = log_probs[b, :, :, blank]
        lpl = np.zeros((Tm, U1m))
        if U > 0:
            idx = np.asarray(labels[b, :U], dtype=np.int64)
            lpl[:, :U] = log_probs[b][:, np.arange(U), idx]
        cost, gb, gl, _, _ = occupancies(lpb, lpl, T, U)
        costs[b] = cost
        grad[b, :, :, blank] -= gb
        if U > 0:
            for u in range(U):
                grad[b, :, u, idx[u]] -= gl[:, u]
    return costs, grad


def joint_logits(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out):
    """rnn_transducer.py:147-156 in fp64.  Returns (h, z): hidden (B,T,U1,J), logits (B,T,U1,V)."""
    e = np.asarray(eouts, np.float64) @ np.asarray(w_enc, np.float64).T + np.asarray(b_enc, np.float64)
    d = np.asarray(douts, np.float64) @ np.asarray(w_dec, np.float64).T + np.asarray(b_dec, np.float64)
    h = np.tanh(e[:, :, None, :] + d[:, None, :, :])
    z = h @ np.asarray(w_out, np.float64).T + np.asarray(b_out, np.float64)
    return e, d, h, z


def log_softmax(z):
    m = z.max(axis=-1, keepdims=True)
    s = z - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True))


def joint_loss_and_grads(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out,
                         labels, tlens, ulens, blank=0):
    """Whole fused path: loss = mean_b cost_b  (reduction="mean", rnn_transducer.py:112) and the
    gradient of that scalar w.r.t. every input of the joint.

    dz[b,t,u,v] = (gamma * softmax(z)[v] - gamma_blank 1[v=blank] - gamma_label 1[v=y_{u+1}]) / B
    (SURVEY.md appendix A), then plain chain rule through Linear/tanh/broadcast-add/Linear.
    """
    eouts = np.asarray(eouts, np.float64)
    douts = np.asarray(douts, np.float64)
    w_out64 = np.asarray(w_out, np.float64)
    e, d, h, z = joint_logits(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out)
    lp = log_softmax(z)
    costs, glp = rnnt_loss_dense(lp, labels, tlens, ulens, blank)
    B = eouts.shape[0]
    loss = costs.mean()
    glp = glp / B
    # log_softmax backward: dz = glp - softmax * sum_v glp
    dz = glp - np.exp(lp) * glp.sum(axis=-1, keepdims=True)
    J = h.shape[-1]
    V = z.shape[-1]
    dz2 = dz.reshape(-1, V)
    d_w_out = dz2.T @ h.reshape(-1, J)
    d_b_out = dz2.sum(axis=0)
    dh = dz @ w_out64
    dpre = dh * (1.0 - h * h)
    d_e = dpre.sum(axis=2)          # (B,T,J)
    d_d = dpre.sum(axis=1)          # (B,U1,J)
    d_w_enc = d_e.reshape(-1, J).T @ eouts.reshape(-1, eouts.shape[-1])
    d_b_enc = d_e.reshape(-1, J).sum(axis=0)
    d_w_dec = d_d.reshape(-1, J).T @ douts.reshape(-1, douts.shape[-1])
    d_b_dec = d_d.reshape(-1, J).sum(axis=0)
    d_eouts = d_e @ np.asarray(w_enc, np.float64)
    d_douts = d_d @ np.asarray(w_dec, np.float64)
    return {
        "loss": loss, "costs": costs, "lp": lp, "dz": dz,
        "d_enc_proj": d_e, "d_dec_proj": d_d,
        "d_eouts": d_eouts, "d_douts": d_douts,
        "d_w_enc": d_w_enc, "d_b_enc": d_b_enc,
        "d_w_dec": d_w_dec, "d_b_dec": d_b_dec,
        "d_w_out": d_w_out, "d_b_out": d_b_out,
    }


def forced_align(log_probs, labels, tlens, ulens, blank=0):
    """RNNTForcedAligner.__call__ (asr/modeling/decoders/rnnt_aligner.py:155-198): best_aligns (B, U1max-1) int32.
    Walk from (0,0); at each cell compare alpha+beta of (t+1,u) and (t,u+1) (:188-196); labels the walk does not
    reach keep the initial 0."""
    B, Tm, U1m, _ = log_probs.shape
    out = np.zeros((B, U1m - 1), dtype=np.int32)
    for b in range(B):
        T, U = int(tlens[b]), int(ulens[b])
        lp = log_probs[b].astype(np.float64)
        lpb = lp[..., blank]
        lpl = np.zeros((Tm, U1m))
        if U > 0:
            lpl[:, :U] = lp[:, np.arange(U), np.asarray(labels[b, :U], dtype=np.int64)]
        alpha, beta, _ = lattice(lpb, lpl, T, U)
        ab = alpha + beta
        t = u = 0
        while t + 1 < T and u < U:
            if ab[t + 1, u] > ab[t, u + 1]:
                t += 1
            else:
                out[b, u] = t
                u += 1
    return out


def word_distill_loss(logits, soft_labels, xlens, ylens):
    """RNNTWordDistillLoss.forward (asr/criteria.py:218-249), normalize_length = normalize_batch = True."""
    loss = 0.0
    for b in range(logits.shape[0]):
        x, y = int(xlens[b]), int(ylens[b])
        lp = log_softmax(logits[b, :x, :y].astype(np.float64))
        loss -= float((soft_labels[b, :y][None] * lp).sum()) / (x * y)
    return loss / logits.shape[0]


def align_distill_loss(logits, soft_labels, aligns, xlens, ylens):
    """RNNTAlignDistillLoss.forward (asr/criteria.py:252-288) AS WRITTEN: `loss_u` is overwritten inside the loop over
    u (:272-278) and subtracted once after it (:280-282), so only the last label's cell contributes."""
    loss = 0.0
    for b in range(logits.shape[0]):
        y = int(ylens[b])
        u = y - 1
        lp = log_softmax(logits[b, int(aligns[b][u]), u].astype(np.float64))
        loss -= float((soft_labels[b, u] * lp).sum()) / y
    return loss / logits.shape[0]

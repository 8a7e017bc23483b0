"""fp64 numpy restatement of the reference's CTC loss path (ORACLE - test infrastructure).

Reference path: asr/modeling/decoders/ctc.py:103-115
    logits = output(eouts)                                     (B,T,V)
    lp     = logits.transpose(1,0).log_softmax(2)              (T,B,V)
    loss   = nn.CTCLoss(blank, reduction="sum", zero_infinity=True)(lp, ys, elens, ylens) / B
The arithmetic of nn.CTCLoss lives in torch ATen (reference pins torch==1.7.1,
asr/correct/README.md:8; image has 2.11).  Its published algorithm is Graves et al. 2006
forward-backward over the blank-extended label sequence; the extension itself is spelled out in
the reference at asr/modeling/decoders/ctc_aligner.py:19-22 (`_label_to_path`).
"""
import numpy as np

NEG_INF = -np.inf


def extend_labels(y, blank):
    """ctc_aligner.py:19-22: l' = [blank, y1, blank, y2, ..., blank], S = 2U+1."""
    U = len(y)
    ext = np.full(2 * U + 1, blank, dtype=np.int64)
    ext[1::2] = y
    return ext


def log_softmax(z):
    m = z.max(axis=-1, keepdims=True)
    s = z - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True))


def ctc_alpha_beta(lp, y, blank):
    """lp (T,V) log-probs of one utterance (already trimmed to T_b), y (U,) labels.
    Returns alpha, beta (T,S) in the "beta includes the emission at t" convention, and nll.
    """
    T = lp.shape[0]
    ext = extend_labels(y, blank)
    S = len(ext)
    lps = lp[:, ext]                                   # (T,S)
    skip = np.zeros(S, dtype=bool)                      # transition s-2 -> s allowed
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    alpha = np.full((T, S), NEG_INF)
    alpha[0, 0] = lps[0, 0]
    if S > 1:
        alpha[0, 1] = lps[0, 1]
    for t in range(1, T):
        a = alpha[t - 1]
        acc = a.copy()
        acc[1:] = np.logaddexp(acc[1:], a[:-1])
        a2 = np.full(S, NEG_INF)
        a2[2:] = a[:-2]
        a2[~skip] = NEG_INF
        acc = np.logaddexp(acc, a2)
        alpha[t] = acc + lps[t]
    if S > 1:
        ll = np.logaddexp(alpha[T - 1, S - 1], alpha[T - 1, S - 2])
    else:
        ll = alpha[T - 1, S - 1]
    beta = np.full((T, S), NEG_INF)
    beta[T - 1, S - 1] = lps[T - 1, S - 1]
    if S > 1:
        beta[T - 1, S - 2] = lps[T - 1, S - 2]
    for t in range(T - 2, -1, -1):
        b = beta[t + 1]
        acc = b.copy()
        acc[:-1] = np.logaddexp(acc[:-1], b[1:])
        b2 = np.full(S, NEG_INF)
        b2[:-2] = b[2:]
        skip_from = np.zeros(S, dtype=bool)             # transition s -> s+2 allowed
        skip_from[:-2] = skip[2:]
        b2[~skip_from] = NEG_INF
        acc = np.logaddexp(acc, b2)
        beta[t] = acc + lps[t]
    return alpha, beta, -ll, ext, lps


def ctc_loss_and_grad(logits, ys, tlens, ulens, blank=0, zero_infinity=True):
    """Whole path ctc.py:109-113 on logits (B,T,V): returns loss (= sum_b nll_b / B), the
    per-utterance nll (after zero_infinity masking) and d loss / d logits (B,T,V).

    grad[b,t,v] = (softmax(z)[v] - occ[t,v]) / B for t < T_b, occ[t,v] =
    sum_{s: l'_s = v} exp(alpha_t(s) + beta_t(s) - lp[t,v] + nll); 0 for padded frames and for
    infeasible utterances (nll = inf) when zero_infinity.
    """
    logits = np.asarray(logits, np.float64)
    B, Tm, V = logits.shape
    nll = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        T, U = int(tlens[b]), int(ulens[b])
        lp = log_softmax(logits[b, :T])
        y = np.asarray(ys[b, :U], dtype=np.int64)
        with np.errstate(invalid="ignore"):
            alpha, beta, nll_b, ext, lps = ctc_alpha_beta(lp, y, blank)
        if not np.isfinite(nll_b):
            nll[b] = 0.0 if zero_infinity else np.inf
            if not zero_infinity:
                grad[b, :T] = np.nan
            continue
        nll[b] = nll_b
        with np.errstate(invalid="ignore"):
            post = np.exp(alpha + beta - lps + nll_b)      # (T,S); -inf - -inf cannot occur at finite lps
        post = np.nan_to_num(post, nan=0.0)
        occ = np.zeros((T, V))
        for s in range(len(ext)):
            occ[:, ext[s]] += post[:, s]
        grad[b, :T] = (np.exp(lp) - occ) / B
    loss = nll.sum() / B
    return loss, nll, grad


def ctc_head_loss_and_grads(eouts, w, bias, ys, tlens, ulens, blank=0):
    """Linear(He,V) head + the loss above (ctc.py:103-113), with grads of loss w.r.t. eouts, w, bias."""
    eouts = np.asarray(eouts, np.float64)
    w = np.asarray(w, np.float64)
    logits = eouts @ w.T + np.asarray(bias, np.float64)
    loss, nll, dz = ctc_loss_and_grad(logits, ys, tlens, ulens, blank)
    V = w.shape[0]
    return {
        "loss": loss, "nll": nll, "logits": logits, "d_logits": dz,
        "d_eouts": dz @ w,
        "d_w": dz.reshape(-1, V).T @ eouts.reshape(-1, eouts.shape[-1]),
        "d_b": dz.reshape(-1, V).sum(axis=0),
    }

"""fp32 numpy restatement of the reference's CTC forced aligner (ORACLE - test infrastructure).

Reference: asr/modeling/decoders/ctc_aligner.py:97-221 (``CTCForcedAligner.__call__``, adapted there from
neural_sp).  What it computes, per utterance b with x = elens[b] valid frames and the blank-extended label path
l' (``_label_to_path``, :19-22) of length P = 2*ylens[b]+1:

  forward   (:172-176)  a_t(s) = lse(a_{t-1}(s), a_{t-1}(s-1), a_{t-1}(s-2) if l'_s != l'_{s-2}) + lp[t, l'_s],
                        a_{-1} = [0, LOG_0, ...]; the lse part is ACCUMULATED into the gathered emissions
                        (``cum_log_prob += log_prob``, :128-129), so that array becomes a_t(s)
  backward  (:178-189)  the same recursion on the flipped path / flipped frames; its lse part
                        B_t(s) = lse over the successors s, s+1, s+2 of (B_{t+1} + emission at t+1) is accumulated
                        on top:  post[t,s] = a_t(s) + B_t(s)
  pick      (:191-219)  greedy, forward in time: among the states reachable from the previously chosen one
                        (same, +1, +2 unless l' repeats; the states whose one-step gamma is not exactly LOG_0) take
                        the argmax of post[t]; every other entry of the row is set to LOG_0 = -1e10 before the argmax
                        (:205-209), so a row whose candidates are all below -1e10 picks the first masked state.

LOG_0 is -1e10, not -inf, and everything is float32: lse(LOG_0 x3) = -1e10 + log 3 == -1e10 and -1e10 + emission
== -1e10 in float32, which is what makes the reference's ``gamma == LOG_0`` test work.  The restatement keeps
float32 and the reference's order of additions.  Frames t >= x never reach the output (:211-213).
"""
import numpy as np

LOG_0 = np.float32(-1e10)


def _lse3(a, b, c):
    m = np.maximum(np.maximum(a, b), c)
    return (m + np.log(np.exp(a - m) + np.exp(b - m) + np.exp(c - m), dtype=np.float32)).astype(np.float32)


def _transition(prev, allow2):
    """lse part of ``_computes_transition`` (:97-125) for one utterance: prev (P,), allow2[s] = l'_s != l'_{s-2}."""
    p1 = np.full_like(prev, LOG_0)
    p2 = np.full_like(prev, LOG_0)
    p1[1:] = prev[:-1]
    p2[2:] = np.where(allow2[2:], prev[:-2], LOG_0)
    return _lse3(prev, p1, p2)


def ctc_forced_align(log_probs, elens, ys, ylens, blank=0):
    """log_probs (B,T,V) log-softmax outputs, elens (B,), ys (B,Umax), ylens (B,) -> best_aligns (B,T) int64,
    zero for t >= elens[b] (ctc_aligner.py:138-221)."""
    lp_all = np.asarray(log_probs, dtype=np.float32)
    B, T, _ = lp_all.shape
    Smax = 2 * ys.shape[1] + 1
    out = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        x, P = int(elens[b]), 2 * int(ylens[b]) + 1
        path = np.full(Smax, blank, dtype=np.int64)
        path[1::2] = ys[b]
        allow2 = np.zeros(Smax, dtype=bool)
        allow2[2:] = path[2:] != path[:-2]
        inside = np.arange(Smax) < P
        em = lp_all[b][:, path]                       # (T,Smax) gathered emissions (:170)
        # ---- forward (:172-176)
        cum = np.empty((x, Smax), dtype=np.float32)
        a = np.full(Smax, LOG_0, dtype=np.float32)
        a[0] = 0.0
        for t in range(x):
            l = _transition(a, allow2)
            l[~inside] = LOG_0
            cum[t] = em[t] + l
            a = l + em[t]
        # ---- backward on the flipped path (:178-189): position k of the flipped path is s = P-1-k
        rpath = path[:P][::-1]
        rallow = np.zeros(P, dtype=bool)
        rallow[2:] = rpath[2:] != rpath[:-2]
        post = np.full((x, Smax), LOG_0, dtype=np.float32)
        bt = np.full(P, LOG_0, dtype=np.float32)
        bt[0] = 0.0
        for t in range(x - 1, -1, -1):
            l = _transition(bt, rallow)
            post[t, :P] = cum[t, :P] + l[::-1]
            bt = l + em[t, :P][::-1]
        post[:, P:] = cum[:, P:] + LOG_0            # outside the path: never reachable below
        # ---- greedy pick (:191-219)
        g = np.full(Smax, LOG_0, dtype=np.float32)
        g[0] = 0.0
        for t in range(x):
            l = _transition(g, allow2)
            l[~inside] = LOG_0
            gam = l + em[t]
            row = np.where(gam == LOG_0, LOG_0, post[t])
            o = int(np.argmax(row))
            out[b, t] = path[o]
            g = np.full(Smax, LOG_0, dtype=np.float32)
            g[o] = 0.0
    return out

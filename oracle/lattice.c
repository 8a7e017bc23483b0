/* Plain-C fp64 restatement of the two lattices on the emoASR sequence-loss path.
 * ORACLE - TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Built by `make -C oracle` into
 * oracle/_build/liboracle_lattice.so and loaded with ctypes from tests/ only.
 *
 * RNN-T: recursion as spelled out in the reference at
 *   asr/modeling/decoders/rnnt_aligner.py:49-83 (alpha) and :121-152 (beta), minus its /T[b];
 *   cost/gradient contract of warp_rnnt.rnnt_loss as called at
 *   asr/modeling/decoders/rnn_transducer.py:106-115.
 * CTC: Graves forward-backward over the blank-extended sequence
 *   (asr/modeling/decoders/ctc_aligner.py:19-22), nn.CTCLoss as configured at
 *   asr/modeling/decoders/ctc.py:36-38.
 */
#include <math.h>
#include <stdlib.h>

static double lae(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log1p(exp(-fabs(a - b)));
}

/* lp2: (T_max, U1_max, 2) doubles = {log p(blank), log p(label u+1)} for one utterance.
 * Writes gamma2 (same layout; zero outside the valid region) and returns cost = -log P(y|x). */
double oracle_rnnt_lattice(const double* lp2, int T_max, int U1_max, int T, int U,
                           double* gamma2) {
    int U1 = U + 1;
    double* alpha = (double*)malloc(sizeof(double) * T * U1);
    double* beta = (double*)malloc(sizeof(double) * T * U1);
#define LPB(t, u) lp2[(((size_t)(t)) * U1_max + (u)) * 2 + 0]
#define LPL(t, u) lp2[(((size_t)(t)) * U1_max + (u)) * 2 + 1]
#define A(t, u) alpha[(size_t)(t) * U1 + (u)]
#define Bt(t, u) beta[(size_t)(t) * U1 + (u)]
    for (int t = 0; t < T; ++t)
        for (int u = 0; u < U1; ++u) {
            if (t == 0 && u == 0) { A(0, 0) = 0.0; continue; }
            double ne = t > 0 ? A(t - 1, u) + LPB(t - 1, u) : -INFINITY;
            double em = u > 0 ? A(t, u - 1) + LPL(t, u - 1) : -INFINITY;
            A(t, u) = lae(ne, em);
        }
    double ll = A(T - 1, U) + LPB(T - 1, U);
    for (int t = T - 1; t >= 0; --t)
        for (int u = U; u >= 0; --u) {
            if (t == T - 1 && u == U) { Bt(t, u) = LPB(t, u); continue; }
            double ne = t < T - 1 ? Bt(t + 1, u) + LPB(t, u) : -INFINITY;
            double em = u < U ? Bt(t, u + 1) + LPL(t, u) : -INFINITY;
            Bt(t, u) = lae(ne, em);
        }
    for (size_t i = 0; i < (size_t)T_max * U1_max * 2; ++i) gamma2[i] = 0.0;
    if (isfinite(ll)) {
        for (int t = 0; t < T; ++t)
            for (int u = 0; u < U1; ++u) {
                double bn = (t < T - 1) ? Bt(t + 1, u) : ((u == U) ? 0.0 : -INFINITY);
                gamma2[(((size_t)t) * U1_max + u) * 2 + 0] = exp(A(t, u) + LPB(t, u) + bn - ll);
                if (u < U)
                    gamma2[(((size_t)t) * U1_max + u) * 2 + 1] =
                        exp(A(t, u) + LPL(t, u) + Bt(t, u + 1) - ll);
            }
    }
    free(alpha);
    free(beta);
    return -ll;
#undef LPB
#undef LPL
#undef A
#undef Bt
}

/* lp: (T, V) log-probs of one utterance; y: U labels.  Writes occ (T, V) = posterior of emitting
 * v at t (zeroed first) and returns nll (may be +inf). */
double oracle_ctc_lattice(const double* lp, int T, int V, const long long* y, int U, int blank,
                          double* occ) {
    int S = 2 * U + 1;
    double* alpha = (double*)malloc(sizeof(double) * T * S);
    double* beta = (double*)malloc(sizeof(double) * T * S);
    int* ext = (int*)malloc(sizeof(int) * S);
    for (int s = 0; s < S; ++s) ext[s] = (s & 1) ? (int)y[s / 2] : blank;
#define LP(t, s) lp[(size_t)(t) * V + ext[s]]
#define A(t, s) alpha[(size_t)(t) * S + (s)]
#define Bt(t, s) beta[(size_t)(t) * S + (s)]
    for (int s = 0; s < S; ++s) A(0, s) = s < 2 ? LP(0, s) : -INFINITY;
    for (int t = 1; t < T; ++t)
        for (int s = 0; s < S; ++s) {
            double a = A(t - 1, s);
            if (s > 0) a = lae(a, A(t - 1, s - 1));
            if (s > 1 && ext[s] != blank && ext[s] != ext[s - 2]) a = lae(a, A(t - 1, s - 2));
            A(t, s) = a + LP(t, s);
        }
    double ll = A(T - 1, S - 1);
    if (S > 1) ll = lae(ll, A(T - 1, S - 2));
    for (int s = 0; s < S; ++s) Bt(T - 1, s) = s >= S - 2 ? LP(T - 1, s) : -INFINITY;
    for (int t = T - 2; t >= 0; --t)
        for (int s = 0; s < S; ++s) {
            double b = Bt(t + 1, s);
            if (s + 1 < S) b = lae(b, Bt(t + 1, s + 1));
            if (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != ext[s]) b = lae(b, Bt(t + 1, s + 2));
            Bt(t, s) = b + LP(t, s);
        }
    for (size_t i = 0; i < (size_t)T * V; ++i) occ[i] = 0.0;
    if (isfinite(ll))
        for (int t = 0; t < T; ++t)
            for (int s = 0; s < S; ++s) {
                double ab = A(t, s) + Bt(t, s);
                if (ab > -INFINITY) occ[(size_t)t * V + ext[s]] += exp(ab - LP(t, s) - ll);
            }
    free(alpha);
    free(beta);
    free(ext);
    return -ll;
#undef LP
#undef A
#undef Bt
}

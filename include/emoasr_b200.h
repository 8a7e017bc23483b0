/* emoasr_b200 -- C ABI of the B200 (sm_100a) sequence-loss hot path of emoASR.
 *
 * Drop-in boundary: the reference is pure Python; the calls that leave it on this path are
 *   warp_rnnt.rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, ...)
 *                          asr/modeling/decoders/rnn_transducer.py:106-115   (third-party CUDA ext)
 *   RNNTDecoder.joint + torch.log_softmax
 *                          asr/modeling/decoders/rnn_transducer.py:101-102, 147-156
 *   nn.CTCLoss(blank, reduction="sum", zero_infinity=True)(log_probs(T,B,V), ys, elens, ylens)
 *                          asr/modeling/decoders/ctc.py:36-38, 109-113, 139-141, 152-154 (ATen)
 * This library is what a maintainer binds instead (ctypes stub in INTEGRATION.md and
 * emoasr_b200/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked host.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises, allocates or reads lengths back to the host.
 *   - the device is the current CUDA device of the calling thread; no global mutable state
 *     besides a thread-local error string, so DataParallel-style threads may call concurrently.
 *   - every function returns 0 on success, an emo_status otherwise; emo_last_error_string()
 *     explains the last failure on the calling thread.
 *   - numerically impossible utterances give cost = +inf (RNN-T, as warp_rnnt) or are zeroed
 *     (CTC with zero_infinity, as torch); kernels never trap.
 *   - lengths: 1 <= tlen[b] <= T, 0 <= ulen[b] <= U1-1.  Cells with t >= tlen[b] or u > ulen[b]
 *     are never read and receive zero gradient.
 */
#ifndef EMOASR_B200_H
#define EMOASR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMO_ABI_VERSION 10

enum emo_status {
    EMO_OK = 0,
    EMO_BAD_ARG = 1,           /* null pointer / non-positive size / misaligned pointer */
    EMO_UNSUPPORTED_SHAPE = 2, /* shape outside what the requested precision path supports */
    EMO_WORKSPACE_TOO_SMALL = 3,
    EMO_LAUNCH_FAILURE = 4,    /* cudaGetLastError() != cudaSuccess after a launch */
    EMO_NO_DEVICE = 5
};

/* arithmetic of the joint's vocabulary projection */
enum emo_precision {
    EMO_PREC_FP32 = 0, /* fp32 FFMA, slab-streamed; parity mode (1e-5 / 1e-4 vs the reference) */
    EMO_PREC_BF16 = 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM; LSE, lattice,
                          occupancies and all reductions stay fp32 */
};

enum emo_op {
    EMO_OP_RNNT_JOINT_FWD = 0,
    EMO_OP_RNNT_JOINT_BWD = 1,
    EMO_OP_CTC = 2,
    EMO_OP_CTC_HEAD = 3,       /* emo_launch_count only: forward + backward of the fused CTC head (J = He) */
    EMO_OP_RNNT_JOINT_FULL = 4 /* emo_launch_count only: emo_rnnt_joint_full_fwd + lattice + _bwd */
};

int emo_abi_version(void);
const char* emo_last_error_string(void); /* host pointer, thread-local storage */

/* Bytes of scratch the op needs for these sizes and precision (host call, no CUDA work). */
size_t emo_workspace_bytes(int op, int precision, int B, int T, int U1, int J, int V);

/* Number of kernel launches one call of the op enqueues for these sizes (host call; lets a harness
 * report how many of this library's kernels ran).  EMO_OP_RNNT_JOINT_FWD counts emo_rnnt_joint_fwd
 * plus the emo_rnnt_lattice_fwd_bwd that follows it; EMO_OP_CTC counts emo_ctc_fwd + emo_ctc_bwd
 * (J is ignored). */
int emo_launch_count(int op, int precision, int B, int T, int U1, int J, int V);

/* 1 if emo_rnnt_joint_fwd / _bwd support these sizes for this precision, else 0.  Host call.  EMO_PREC_FP32
 * supports every shape; EMO_PREC_BF16 needs J % 128 == 0, J <= 512, B <= 1024, T and U1 < 65536 (any vocabulary
 * size: one that is not a multiple of 32 is padded inside the workspace, zero weights / -1e30 bias). */
int emo_rnnt_joint_supported(int precision, int B, int T, int U1, int J, int V);

/* ---- RNN-T lattice on gathered pairs ---------------------------------------------------------
 * Replaces the alpha/beta/grad kernels of warp_rnnt.rnnt_loss (rnn_transducer.py:106-115).
 * lp2      (B,T,U1,2)  {log p(blank | t,u), log p(y_{u+1} | t,u)}
 * tlen,ulen (B)        int32 (the reference casts with .int(), rnn_transducer.py:108-110)
 * alpha_ws, beta_ws (B,T,U1) scratch, overwritten
 * cost     (B)         -log P(y_b | x_b)                       (no reduction, no normalisation)
 * gamma2   (B,T,U1,2)  {-d cost/d lp_blank, -d cost/d lp_label} = transition posteriors; zero
 *                      outside the valid region.
 */
int emo_rnnt_lattice_fwd_bwd(const float* lp2, const int* tlen, const int* ulen,
                             int B, int T, int U1,
                             float* alpha_ws, float* beta_ws,
                             float* cost, float* gamma2, void* stream);

/* ---- forced alignment on the lattice ----------------------------------------------------------
 * Replaces RNNTForcedAligner.__call__ (asr/modeling/decoders/rnnt_aligner.py:155-198: two Numba spin-lock kernels
 * + a Python walk per utterance): given alpha_ws / beta_ws of emo_rnnt_lattice_fwd_bwd (or emo_rnnt_dense_fwd),
 * aligns (B,U1-1) int32: aligns[b,u] = frame at which y_u is emitted on the greedy walk through alpha + beta
 * (0 for labels the walk does not reach -- the reference's behaviour). */
int emo_rnnt_align(const float* alpha_ws, const float* beta_ws, const int* tlen, const int* ulen, int B, int T,
                   int U1, int* aligns, void* stream);

/* ---- warp_rnnt module seam: dense log-probs in, sparse gradient out --------------------------
 * Same contract as warp_rnnt.rnnt_loss(log_probs, labels, frames_lengths, labels_lengths,
 * average_frames=False, reduction=None, blank, gather=False).
 * log_probs (B,T,U1,V) fp32 contiguous; labels (B,U1-1) int32.
 * lp2_ws/gamma2_ws (B,T,U1,2), alpha_ws/beta_ws (B,T,U1) scratch.
 * emo_rnnt_dense_fwd gathers, runs the lattice and leaves gamma2_ws for the backward call.
 * emo_rnnt_dense_bwd writes grad_log_probs (B,T,U1,V) = -grad_cost[b] * gamma at the blank and
 * label entries and 0 elsewhere (the whole tensor is written). */
int emo_rnnt_dense_fwd(const float* log_probs, const int* labels, const int* tlen, const int* ulen,
                       int B, int T, int U1, int V, int blank,
                       float* lp2_ws, float* alpha_ws, float* beta_ws,
                       float* cost, float* gamma2_ws, void* stream);
int emo_rnnt_dense_bwd(const float* gamma2_ws, const int* labels, const int* tlen, const int* ulen,
                       const float* grad_cost, int B, int T, int U1, int V, int blank,
                       float* grad_log_probs, void* stream);

/* ---- fused joint (rnn_transducer.py:147-156 + :102) -----------------------------------------
 * enc_proj (B,T,J)  = w_enc(eouts) + b_enc      dec_proj (B,U1,J) = w_dec(douts) + b_dec
 * w_out (V,J) fp32 row-major (torch Linear weight), b_out (V)
 * For every valid cell: h = tanh(enc_proj[b,t] + dec_proj[b,u]); z = w_out h + b_out;
 *   lse[b,t,u] = logsumexp_v z;  lp2[b,t,u] = {z[blank]-lse, z[labels[b,u]]-lse (u < ulen[b])}.
 * EMO_PREC_BF16: nothing of size N x V (N = valid lattice cells) is written to memory by the forward, and the
 * backward keeps it that way (see emo_rnnt_joint_bwd); the only per-cell outputs are lp2 and lse.
 * EMO_PREC_FP32 streams the logits through a bounded slab inside `ws`.
 */
int emo_rnnt_joint_fwd(const float* enc_proj, const float* dec_proj,
                       const float* w_out, const float* b_out,
                       const int* labels, const int* tlen, const int* ulen,
                       int B, int T, int U1, int J, int V, int blank, int precision,
                       float* lp2, float* lse, void* ws, size_t ws_bytes, void* stream);

/* Backward of cost (B) w.r.t. enc_proj, dec_proj, w_out, b_out given grad_cost (B):
 *   dz[b,t,u,v] = grad_cost[b] * ((g_blank+g_label) * exp(z[v]-lse) - g_blank 1[v=blank]
 *                                  - g_label 1[v=labels[b,u]])
 *   d_w_out (V,J) = sum dz^T h ; d_b_out (V) = sum dz ; dh = dz w_out ;
 *   dpre = dh (1-h^2) ; d_enc_proj[b,t] = sum_u dpre ; d_dec_proj[b,u] = sum_t dpre.
 * EMO_PREC_BF16: z is RECOMPUTED tile by tile on the tensor cores by producer CTA pairs of one persistent
 * kernel, turned into dz in their epilogue and handed to the dh / dW consumer pairs of the same kernel through
 * a ring of tiles inside `ws` whose size (36 MB) does not depend on the problem size and stays L2-resident; dz
 * is never a tensor in HBM.  What does go through HBM is dh as bf16 (N x J), read back once by the axis
 * reductions.
 * grad_lse (B,T,U1), optional (NULL = none): gradient w.r.t. the forward's `lse` output, for losses that use it
 * next to the transducer cost (the distillation losses of asr/criteria.py:218-288 need sum_v q log_softmax(z) =
 * q.z - (sum q) lse); it adds grad_lse * softmax(z) to dz.  Must be 0 for cells outside the valid lattice.
 * All four outputs are overwritten (not accumulated into).  enc_proj, dec_proj, w_out, b_out, labels, the
 * lengths, lse and lp2 must be the tensors the forward call saw / produced (as autograd's saved tensors are):
 * the backward re-reads them (h is recomputed from enc_proj / dec_proj; lp2 gives the exact blank / label
 * probabilities). */
int emo_rnnt_joint_bwd(const float* enc_proj, const float* dec_proj,
                       const float* w_out, const float* b_out,
                       const int* labels, const int* tlen, const int* ulen,
                       const float* lse, const float* lp2, const float* gamma2, const float* grad_cost,
                       const float* grad_lse,
                       int B, int T, int U1, int J, int V, int blank, int precision,
                       float* d_enc_proj, float* d_dec_proj, float* d_w_out, float* d_b_out,
                       void* ws, size_t ws_bytes, void* stream);

/* ---- fused joint from the encoder / prediction-network outputs (projections folded in) -----------
 * emo_rnnt_joint_fwd / _bwd plus `enc_proj = w_enc(eouts) + b_enc`, `dec_proj = w_dec(douts) + b_dec`
 * (rnn_transducer.py:57-58,153) and their backward, tensor-core mode only.  The projected streams exist only as the
 * fp16 copies the joint kernels gather from.  eouts (B,T,He), douts (B,U1,Hd), w_enc (J,He), w_dec (J,Hd) fp32.
 * The forward's workspace `fws` (emo_rnnt_joint_full_workspace_bytes(0, ...), 256-byte aligned) holds the bf16 / fp16
 * operand copies and must be handed UNCHANGED to the backward, which reads them instead of casting again; the
 * backward's own workspace is op 1.  All eight gradient outputs are overwritten.  He, Hd multiples of 16; J, B, T, U1
 * as emo_rnnt_joint_supported. */
int emo_rnnt_joint_full_supported(int B, int T, int U1, int He, int Hd, int J, int V);
size_t emo_rnnt_joint_full_workspace_bytes(int op, int B, int T, int U1, int He, int Hd, int J, int V);
int emo_rnnt_joint_full_fwd(const float* eouts, const float* douts, const float* w_enc, const float* b_enc,
                            const float* w_dec, const float* b_dec, const float* w_out, const float* b_out,
                            const int* labels, const int* tlen, const int* ulen, int B, int T, int U1, int He, int Hd,
                            int J, int V, int blank, float* lp2, float* lse, void* fws, size_t fws_bytes, void* stream);
int emo_rnnt_joint_full_bwd(const float* b_out, const int* labels, const int* tlen, const int* ulen, const float* lse,
                            const float* lp2, const float* gamma2, const float* grad_cost, const float* grad_lse,
                            const void* fws, int B, int T, int U1, int He, int Hd, int J, int V, int blank,
                            float* d_eouts, float* d_douts, float* d_w_enc, float* d_b_enc, float* d_w_dec,
                            float* d_b_dec, float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, void* stream);

/* ---- CTC (ctc.py:109-113; torch.nn.CTCLoss semantics) ----------------------------------------
 * logits (B,T,V) fp32 contiguous -- the log_softmax of ctc.py:110 is fused: the kernels read raw
 * logits.  labels (B,Umax) int64 padded arbitrarily; tlen, ulen (B) int64 (as the reference
 * passes them).  S = 2*Umax+1.
 * emo_ctc_fwd:  lse (B,T), alpha_ws (B,T,S) scratch kept for backward,
 *               beta_ws (B,T,S) or NULL: when given, the beta lattice runs side by side with alpha
 *               (what a training step wants); NULL = forward only.
 *               nll (B): -log P(y|x), or 0 when infeasible and zero_infinity != 0 (else +inf).
 * emo_ctc_bwd:  grad_logits (B,T,V) = grad_nll[b] * (softmax(z)[t,v] - occ[t,v]) for t < tlen[b],
 *               0 for padded frames and for infeasible utterances (whole tensor written).
 *               beta_ws (B,T,S): beta from emo_ctc_fwd (beta_valid != 0) or scratch the beta lattice
 *               is computed into first (beta_valid == 0).
 */
int emo_ctc_fwd(const float* logits, const long long* labels, const long long* tlen,
                const long long* ulen, int B, int T, int V, int Umax, int blank, int zero_infinity,
                float* lse, float* alpha_ws, float* beta_ws, float* nll, void* stream);
int emo_ctc_bwd(const float* logits, const long long* labels, const long long* tlen,
                const long long* ulen, const float* lse, const float* alpha_ws, const float* nll,
                const float* grad_nll, int B, int T, int V, int Umax, int blank, int zero_infinity,
                float* beta_ws, int beta_valid, float* grad_logits, void* stream);

/* ---- CTC forced alignment ---------------------------------------------------------------------------
 * Replaces CTCForcedAligner.__call__ (asr/modeling/decoders/ctc_aligner.py:138-221; used by the CTC distillation
 * path, ctc.py:117-120,158-161): forward / backward over the blank-extended label path and the reference's greedy
 * pick through the state posteriors, one launch for the batch instead of three Python loops over the frames with a
 * host read per frame.  log_probs (B,T,V) fp32 = log_softmax(logits) as the reference passes it (NOT modified: the
 * reference zeroes the padded frames of its argument in place, :144-147; they never reach the result); labels /
 * lengths as emo_ctc_fwd; aligns (B,T) int64: the label (or blank) chosen for every frame t < tlen[b], 0 after it.
 * ws: emo_ctc_align_workspace_bytes(B,T,Umax) bytes.  2*Umax+1 <= 1024, Umax >= 1. */
size_t emo_ctc_align_workspace_bytes(int B, int T, int Umax);
int emo_ctc_align(const float* log_probs, const long long* labels, const long long* tlen, const long long* ulen,
                  int B, int T, int V, int Umax, int blank, long long* aligns, void* ws, size_t ws_bytes,
                  void* stream);

/* ---- fused CTC head: output Linear + log_softmax + CTC loss --------------------------------------
 * Replaces asr/modeling/decoders/ctc.py:103-113 (`logits = self.output(eouts)`; `ctc_loss_fn(logits.transpose(1,0)
 * .log_softmax(2), ys, elens, ylens)`) including the Linear's backward: the (B,T,V) logits, log-probs and their
 * gradient are never written to memory.
 * eouts (B,T,He) fp32, w (V,He) fp32 row-major (torch Linear weight), b (V); labels / lengths as emo_ctc_fwd.
 * Tensor-core path (operands rounded to bf16, fp32 accumulation; loss <= 2e-3, gradients <= 2e-2 relative):
 * needs He % 128 == 0, He <= 512, 2*Umax+1 <= 1024, B <= 1024 and a vocabulary of at most ~16 k entries
 * (emo_ctc_head_supported; any V is padded to a multiple of 32 inside the workspace).
 * emo_ctc_head_fwd: lse (B,T) log-sum-exp of every valid frame's logits; emis (B,T,Umax+1) emission log-probs
 *                   {blank, y_0 .. y_{U_b-1}} per frame (kept for the backward); alpha_ws / beta_ws (B,T,2*Umax+1)
 *                   lattices (beta_ws may be NULL when no gradient will be asked for); nll (B).
 * emo_ctc_head_bwd: d_eouts (B,T,He), d_w (V,He), d_b (V) of sum_b grad_nll[b] * nll[b]; all three are overwritten.
 *                   lse / emis / alpha_ws / beta_ws must be the forward's outputs.
 * emo_ctc_head_workspace_bytes: op 0 = forward, 1 = backward; 0 for unsupported shapes.  Host calls. */
int emo_ctc_head_supported(int B, int T, int He, int V, int Umax);
size_t emo_ctc_head_workspace_bytes(int op, int B, int T, int He, int V, int Umax);
int emo_ctc_head_fwd(const float* eouts, const float* w, const float* b, const long long* labels,
                     const long long* tlen, const long long* ulen, int B, int T, int He, int V, int Umax,
                     int blank, int zero_infinity, float* lse, float* emis, float* alpha_ws, float* beta_ws,
                     float* nll, void* ws, size_t ws_bytes, void* stream);
int emo_ctc_head_bwd(const float* eouts, const float* w, const float* b, const long long* labels,
                     const long long* tlen, const long long* ulen, const float* lse, const float* emis,
                     const float* alpha_ws, const float* beta_ws, const float* grad_nll, int B, int T, int He,
                     int V, int Umax, int blank, float* d_eouts, float* d_w, float* d_b, void* ws, size_t ws_bytes,
                     void* stream);

/* ---- decode-time joint -------------------------------------------------------------------------------
 * Replaces the (1,1,.) joint calls of the reference's search loops (rnn_transducer.py:194-325: `self.joint(eouts[b:b+1,
 * t:t+1], dout)` -> argmax / log_softmax per step).  fp32 arithmetic (hypotheses should match the reference's).
 *   z[n,:] = w_out tanh(enc_proj[row_n,:] + dec_proj[n,:]) + b_out,  n < N (utterances of a batch / hypotheses of a beam)
 * enc_proj (rows,J) = w_enc(eouts)+b flattened over (b,t); enc_row (N) int32 row per n, or NULL (row n);
 * dec_proj (N,J) = w_dec(dout)+b; J % 4 == 0.  ws: emo_rnnt_step_workspace_bytes(N) bytes, ZERO before its first use
 * (the calls leave it zeroed), 256-byte aligned.
 * emo_rnnt_joint_step:  logits (N,V) and / or token (N) = argmax_v z (lowest index on ties); either may be NULL.
 * emo_rnnt_greedy_step: one step of batched greedy search (rnn_transducer.py:194-240) for all N utterances at once,
 *   enc_row = n * T + t_idx[n]: token[n] = argmax; rows with t_idx < tlen and hyp_len <= max_len are active: the token
 *   is appended to align (N,align_cap; optional), then blank -> t_idx += 1, else hyp[n][hyp_len++] = token (hyp is
 *   (N,max_len+1)) and emitted[n] = 1 (the caller then advances that row's prediction network);
 *   n_active (1, optional) = rows still active after the step.  No host synchronisation. */
size_t emo_rnnt_step_workspace_bytes(int N);
int emo_rnnt_joint_step(const float* enc_proj, const int* enc_row, const float* dec_proj, const float* w_out,
                        const float* b_out, int N, int J, int V, float* logits, long long* token, void* ws,
                        size_t ws_bytes, void* stream);
int emo_rnnt_greedy_step(const float* enc_proj, const float* dec_proj, const float* w_out, const float* b_out,
                         const int* tlen, int N, int T, int J, int V, int blank, int max_len, int* t_idx, int* hyp,
                         int* hyp_len, int* align, int* align_len, int align_cap, unsigned char* emitted,
                         long long* token, int* n_active, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMOASR_B200_H */

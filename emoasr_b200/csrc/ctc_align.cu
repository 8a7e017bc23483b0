// CTC forced alignment on the device (include/emoasr_b200.h: emo_ctc_align).
// Replaces CTCForcedAligner.__call__ (asr/modeling/decoders/ctc_aligner.py:138-221): there the three passes are
// Python loops over the frames with O(10) tensor ops each and, in the pick pass, a per-frame argmax read back to the
// host for every utterance.  Here: one CTA per utterance, thread == position s of the blank-extended label path,
// two sweeps over the frames inside one launch:
//   sweep 1 (t = x-1 .. 0)  B_t(s) = lse over the successors s, s+1, s+2 of (B_{t+1} + emission at t+1)  -> workspace
//   sweep 2 (t = 0 .. x-1)  a_t(s) = lse over the predecessors of a_{t-1} (+ emission), post = a_t(s) + B_t(s), and
//                           the reference's greedy pick: among the states reachable from the previously chosen one
//                           the argmax of post; all other states count as LOG_0 = -1e10 (ctc_aligner.py:205-209).
// Arithmetic follows the reference's float32 order of operations (LOG_0 is -1e10, not -inf: "unreachable" is detected
// there by equality with it), so the picks agree with it except on exact floating-point ties.
#include "common.cuh"

namespace emo {
namespace {

constexpr float kLog0 = -1e10f;
constexpr int kBlk = 8;   // frames whose emissions are fetched ahead of their use

__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__global__ void ctc_align_kernel(const float* __restrict__ lp, const long long* __restrict__ labels,
                                 const long long* __restrict__ tlen, const long long* __restrict__ ulen, int T,
                                 int V, int Umax, int blank, float* __restrict__ ws, long long* __restrict__ aligns) {
    extern __shared__ float sh[];
    const int Smax = 2 * Umax + 1;
    float* v0 = sh;                       // [Smax + 4]  recursion values, double-buffered, 2 guard cells each side
    float* v1 = v0 + Smax + 4;
    float* cand = v1 + Smax + 4;          // [4] post of the states s = o, o+1, o+2
    int* path = reinterpret_cast<int*>(cand + 4);   // [Smax + 2]
    const int b = blockIdx.x, s = threadIdx.x;
    const int x = (int)min(max(tlen[b], 0LL), (long long)T);
    const int P = 2 * (int)min(max(ulen[b], 0LL), (long long)Umax) + 1;
    for (int i = s; i < Smax + 2; i += blockDim.x) {
        int l = blank;
        if (i < Smax && (i & 1)) l = (int)min(max(labels[(size_t)b * Umax + (i >> 1)], 0LL), (long long)(V - 1));
        path[i] = i < Smax ? l : -1;
    }
    for (int i = s; i < 2 * (Smax + 4) + 4; i += blockDim.x) sh[i] = kLog0;
    long long* out = aligns + (size_t)b * T;
    for (int t = x + s; t < T; t += blockDim.x) out[t] = 0;          // ctc_aligner.py:192: zeros past the utterance
    __syncthreads();
    const bool in = s < P;
    const int lab = s < Smax ? path[s] : 0;
    const bool skip_fwd = s >= 2 && s < Smax && path[s] != path[s - 2];          // s-2 -> s allowed
    const bool skip_bwd = s + 2 < Smax && path[s + 2] != path[s];                // s -> s+2 allowed
    const float* lpb = lp + (size_t)b * T * V + lab;
    float* wsb = ws + (size_t)b * T * Smax + s;
    float ec[kBlk], en[kBlk];

    // ---------------- sweep 1: B_t(s), t descending.  bt(s) = B_{t+1}(s) + em_{t+1}(s); before the last frame it is
    // 0 at the last path position and LOG_0 elsewhere (the flipped [LOG_1, LOG_0, ...] of ctc_aligner.py:160-163)
    float* cur = v0 + 2;
    float* nxt = v1 + 2;
    if (s == P - 1) cur[s] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kBlk; ++i) {
        const int t = x - 1 - i;
        ec[i] = (in && t >= 0) ? __ldg(lpb + (size_t)t * V) : 0.f;
    }
    for (int tb = x - 1; tb >= 0; tb -= kBlk) {
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            const int t = tb - kBlk - i;
            en[i] = (in && t >= 0) ? __ldg(lpb + (size_t)t * V) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            const int t = tb - i;
            if (t >= 0) {                                        // block-uniform
                if (in) {
                    const float a0 = cur[s];
                    const float a1 = s + 1 < P ? cur[s + 1] : kLog0;
                    const float a2 = (s + 2 < P && skip_bwd) ? cur[s + 2] : kLog0;
                    const float l = lse3(a0, a1, a2);
                    wsb[(size_t)t * Smax] = l;
                    nxt[s] = l + ec[i];
                }
                __syncthreads();
                float* tmp = cur; cur = nxt; nxt = tmp;
            }
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) ec[i] = en[i];
    }

    // ---------------- sweep 2: a_t(s), post, greedy pick
    __syncthreads();
    for (int i = s; i < Smax + 4; i += blockDim.x) { v0[i] = kLog0; v1[i] = kLog0; }
    __syncthreads();
    cur = v0 + 2;
    nxt = v1 + 2;
    if (s == 0) cur[0] = 0.f;                                    // alpha before the first frame (:160-161)
    __syncthreads();
    float wc[kBlk], wn[kBlk];
#pragma unroll
    for (int i = 0; i < kBlk; ++i) {
        ec[i] = (in && i < x) ? __ldg(lpb + (size_t)i * V) : 0.f;
        wc[i] = (in && i < x) ? wsb[(size_t)i * Smax] : 0.f;
    }
    int o = 0;                                                   // gamma before the first frame: state 0 (:163)
    for (int tb = 0; tb < x; tb += kBlk) {
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            const int t = tb + kBlk + i;
            en[i] = (in && t < x) ? __ldg(lpb + (size_t)t * V) : 0.f;
            wn[i] = (in && t < x) ? wsb[(size_t)t * Smax] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) {
            const int t = tb + i;
            if (t < x) {                                         // block-uniform
                if (in) {
                    const float a0 = cur[s];
                    const float a1 = s >= 1 ? cur[s - 1] : kLog0;
                    const float a2 = skip_fwd ? cur[s - 2] : kLog0;
                    const float l = lse3(a0, a1, a2);
                    const float post = (ec[i] + l) + wc[i];      // cum_log_prob += alpha part, then += beta part
                    nxt[s] = l + ec[i];
                    if (s >= o && s <= o + 2) cand[s - o] = post;
                }
                __syncthreads();
                // every thread takes the same decision from the three candidates (ctc_aligner.py:195-219)
                float best = -INFINITY;
                int pick = -1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int sc = o + c;
                    const bool ok = sc < P && (c < 2 || path[sc] != path[sc - 2]);
                    const float v = ok ? cand[c] : -INFINITY;
                    if (ok && v > best) { best = v; pick = sc; }
                }
                if (!(best > kLog0)) {
                    // no candidate above LOG_0: the reference's argmax sees LOG_0 at every masked state and takes
                    // the first maximum of the whole row
                    const float vmax = fmaxf(best, kLog0);
                    pick = 0;
                    for (int sc = 0; sc < Smax; ++sc) {
                        const int c = sc - o;
                        const bool ok = c >= 0 && c <= 2 && sc < P && (c < 2 || path[sc] != path[sc - 2]);
                        const float v = ok ? cand[c] : kLog0;
                        if (v == vmax) { pick = sc; break; }
                    }
                }
                o = pick;
                if (s == 0) out[t] = path[o];
                __syncthreads();                                 // cand is rewritten in the next step
                float* tmp = cur; cur = nxt; nxt = tmp;
            }
        }
#pragma unroll
        for (int i = 0; i < kBlk; ++i) { ec[i] = en[i]; wc[i] = wn[i]; }
    }
}

}  // namespace
}  // namespace emo

using namespace emo;

extern "C" size_t emo_ctc_align_workspace_bytes(int B, int T, int Umax) {
    if (B <= 0 || T <= 0 || Umax < 0) return 0;
    return (size_t)B * T * (2 * Umax + 1) * sizeof(float);
}

extern "C" int emo_ctc_align(const float* log_probs, const long long* labels, const long long* tlen,
                             const long long* ulen, int B, int T, int V, int Umax, int blank, long long* aligns,
                             void* ws, size_t ws_bytes, void* stream) {
    EMO_REQUIRE(log_probs && labels && tlen && ulen && aligns && ws, EMO_BAD_ARG, "ctc_align: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && V > 0 && Umax >= 1 && blank >= 0 && blank < V, EMO_BAD_ARG,
                "ctc_align: bad sizes B=%d T=%d V=%d Umax=%d blank=%d", B, T, V, Umax, blank);
    const int Smax = 2 * Umax + 1;
    EMO_REQUIRE(Smax <= 1024, EMO_UNSUPPORTED_SHAPE, "ctc_align: 2*Umax+1 = %d > 1024", Smax);
    EMO_REQUIRE(ws_bytes >= emo_ctc_align_workspace_bytes(B, T, Umax), EMO_BAD_ARG, "ctc_align: workspace too small");
    const int threads = (Smax + 31) / 32 * 32;
    const size_t smem = (2 * (Smax + 4) + 4) * sizeof(float) + (Smax + 2) * sizeof(int);
    ctc_align_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(log_probs, labels, tlen, ulen, T, V, Umax, blank,
                                                                  (float*)ws, aligns);
    EMO_CHECK_LAUNCH("ctc_align_kernel");
    return EMO_OK;
}

// Decode-time joint: one lattice cell per row instead of the whole (T, U+1) plane.
//
//   z[n,:] = w_out tanh(enc_proj[row_n,:] + dec_proj[n,:]) + b_out           (rnn_transducer.py:147-156 at T = L = 1)
//
// The reference decodes with (1,1,.) cuBLAS calls in Python loops (rnn_transducer.py:194-325): three Linear
// launches, a broadcast add, a tanh and an argmax + .item() per emitted symbol or consumed frame, one utterance at a
// time.  Here one launch does the joint for N rows (N = utterances of a batch in greedy search, hypotheses of a beam
// in ALSD), the argmax over the vocabulary and -- for greedy search -- the whole bookkeeping of the step (advance the
// frame on blank, append the token otherwise), so the host never reads a token back inside the loop.
//
// fp32 FFMA on purpose: the search compares logits, and hypotheses should match the reference's fp32 arithmetic.
// The weight matrix (V x J fp32, 2 MB at cfg 3) is L2-resident across steps; the kernel is latency-bound (~5 us).
//   block   8 warps; h = tanh(enc + dec) of up to 32 rows in shared memory (recomputed per block: N * J tanh)
//   warp    one vocabulary row at a time: lanes stride over J with float4 loads, 32 row accumulators per lane,
//           butterfly reduce-scatter (31 shuffles) leaves z[n] in lane n
//   argmax  per lane across the warp's vocabulary rows, then per block through shared memory, then one 64-bit
//           atomicMax per (block, row) on a key (ordered value << 32 | ~index): largest value, lowest index wins
//   last block (atomic ticket) turns the keys into tokens and applies the greedy update.
#include "common.cuh"

namespace emo {
namespace {

constexpr int kStepWarps = 8;
constexpr int kStepRows = 32;    // rows (utterances / hypotheses) per pass

struct StepArgs {
    const float* enc;        // (rows_enc, J)
    const int* enc_row;      // (N) row of enc for every n, or NULL: n-th row
    const float* dec;        // (N, J)
    const float* w;          // (V, J)
    const float* b;          // (V)
    float* logits;           // (N, V) or NULL
    unsigned long long* key; // (N) workspace, zeroed by the caller of the kernel
    unsigned int* ticket;    // (1) workspace, zeroed
    long long* token;        // (N) argmax, or NULL
    int N, J, V;
    // greedy bookkeeping (all NULL / 0 when not used)
    int* t_idx;              // (N) current frame
    const int* tlen;         // (N)
    int T;                   // frames per utterance in enc: enc_row = n * T + t_idx[n]
    int* hyp;                // (N, max_len + 1)
    int* hyp_len;            // (N)
    int* align;              // (N, align_cap) or NULL
    int* align_len;          // (N)
    unsigned char* emitted;  // (N) 1 if the row appended a token in this step
    int* n_active;           // (1) rows that are still decoding after this step
    int blank, max_len, align_cap;
};

__device__ __forceinline__ unsigned int ordered(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// greedy: a row is active while it has frames left and its hypothesis is not over-long (rnn_transducer.py:218-233)
__device__ __forceinline__ bool row_active(const StepArgs& a, int n) {
    return a.t_idx[n] < a.tlen[n] && a.hyp_len[n] <= a.max_len;
}

__global__ void __launch_bounds__(kStepWarps * 32) joint_step_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float s_h[];                 // [kStepRows][J]
    __shared__ unsigned long long s_best[kStepWarps][kStepRows];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int J4 = a.J >> 2;
    for (int n0 = 0; n0 < a.N; n0 += kStepRows) {
        const int nb = min(kStepRows, a.N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * J4; i += blockDim.x) {
            const int n = n0 + i / J4, j4 = i % J4;
            int er = n;
            if (a.t_idx) er = n * a.T + min(a.t_idx[n], a.T - 1);
            else if (a.enc_row) er = a.enc_row[n];
            const float4 e = __ldg(reinterpret_cast<const float4*>(a.enc + (size_t)er * a.J) + j4);
            const float4 d = __ldg(reinterpret_cast<const float4*>(a.dec + (size_t)n * a.J) + j4);
            reinterpret_cast<float4*>(s_h)[i] = make_float4(tanhf(e.x + d.x), tanhf(e.y + d.y), tanhf(e.z + d.z),
                                                            tanhf(e.w + d.w));
        }
        __syncthreads();
        unsigned long long best = 0ull;
        for (int v = blockIdx.x * kStepWarps + warp; v < a.V; v += gridDim.x * kStepWarps) {
            float acc[kStepRows];
#pragma unroll
            for (int n = 0; n < kStepRows; ++n) acc[n] = 0.f;
            const float4* wrow = reinterpret_cast<const float4*>(a.w + (size_t)v * a.J);
            for (int j4 = lane; j4 < J4; j4 += 32) {
                const float4 w4 = __ldg(wrow + j4);
#pragma unroll
                for (int n = 0; n < kStepRows; ++n) {
                    if (n < nb) {
                        const float4 h4 = reinterpret_cast<const float4*>(s_h)[n * J4 + j4];
                        acc[n] = fmaf(w4.x, h4.x, fmaf(w4.y, h4.y, fmaf(w4.z, h4.z, fmaf(w4.w, h4.w, acc[n]))));
                    }
                }
            }
            // butterfly reduce-scatter: after the step with offset o a lane keeps the o accumulators whose index has
            // the same `o` bit as the lane; 16 + 8 + 4 + 2 + 1 shuffles, lane n ends with the total of row n
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const bool hi = lane & o;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const float send = hi ? acc[i] : acc[i + o];
                    const float keep = hi ? acc[i + o] : acc[i];
                    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            if (lane < nb) {
                const float z = acc[0] + __ldg(a.b + v);
                if (a.logits) a.logits[(size_t)(n0 + lane) * a.V + v] = z;
                const unsigned long long k = ((unsigned long long)ordered(z) << 32) | (unsigned int)(0xffffffffu - (unsigned int)v);
                best = k > best ? k : best;
            }
        }
        if (a.key) {
            s_best[warp][lane] = best;
            __syncthreads();
            if (warp == 0 && lane < nb) {
                unsigned long long m = s_best[0][lane];
#pragma unroll
                for (int w = 1; w < kStepWarps; ++w) m = s_best[w][lane] > m ? s_best[w][lane] : m;
                if (m) atomicMax(a.key + n0 + lane, m);
            }
        }
    }
    if (!a.key) return;
    // ---- last block: keys -> tokens, greedy bookkeeping, workspace reset for the next step
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int active = 0;
    for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
        const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(a.key + n);
        const int tok = (int)(0xffffffffu - (unsigned int)(k & 0xffffffffu));
        a.key[n] = 0ull;
        if (a.token) a.token[n] = tok;
        if (!a.t_idx) continue;
        unsigned char em = 0;
        if (row_active(a, n)) {
            if (a.align && a.align_len[n] < a.align_cap) a.align[(size_t)n * a.align_cap + a.align_len[n]] = tok;
            if (a.align_len) a.align_len[n] += 1;
            if (tok == a.blank) {
                a.t_idx[n] += 1;
            } else {
                a.hyp[(size_t)n * (a.max_len + 1) + a.hyp_len[n]] = tok;
                a.hyp_len[n] += 1;
                em = 1;
            }
            active += row_active(a, n) ? 1 : 0;
        }
        a.emitted[n] = em;
    }
    if (a.n_active) {
        __shared__ int s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (active) atomicAdd(&s_cnt, active);
        __syncthreads();
        if (threadIdx.x == 0) *a.n_active = s_cnt;
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
}

int step_launch(StepArgs& a, void* ws, size_t ws_bytes, bool needs_key, cudaStream_t st) {
    EMO_REQUIRE(a.enc && a.dec && a.w && a.b, EMO_BAD_ARG, "joint_step: null pointer");
    EMO_REQUIRE(a.N > 0 && a.J > 0 && a.V > 0, EMO_BAD_ARG, "joint_step: bad sizes");
    EMO_REQUIRE(a.J % 4 == 0 && ((uintptr_t)a.enc & 15) == 0 && ((uintptr_t)a.dec & 15) == 0 && ((uintptr_t)a.w & 15) == 0,
                EMO_UNSUPPORTED_SHAPE, "joint_step: J must be a multiple of 4 and the pointers 16-byte aligned");
    const size_t smem = (size_t)kStepRows * a.J * sizeof(float);
    EMO_REQUIRE(smem <= 200 * 1024, EMO_UNSUPPORTED_SHAPE, "joint_step: J = %d too large", a.J);
    a.key = nullptr;
    a.ticket = nullptr;
    if (needs_key) {
        const size_t need = align_up((size_t)a.N * sizeof(unsigned long long), 256) + 256;
        EMO_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, EMO_WORKSPACE_TOO_SMALL,
                    "joint_step: workspace too small (%zu bytes needed) or not 256-byte aligned", need);
        a.key = reinterpret_cast<unsigned long long*>(ws);
        a.ticket = reinterpret_cast<unsigned int*>((char*)ws + align_up((size_t)a.N * sizeof(unsigned long long), 256));
    }
    if (smem > 48 * 1024)
        EMO_CUDA(cudaFuncSetAttribute(joint_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = min(ceil_div(a.V, kStepWarps), 2 * sm_count());
    joint_step_kernel<<<grid, kStepWarps * 32, smem, st>>>(a);
    EMO_CHECK_LAUNCH("joint_step_kernel");
    return EMO_OK;
}

}  // namespace
}  // namespace emo

using namespace emo;

extern "C" size_t emo_rnnt_step_workspace_bytes(int N) {
    return N > 0 ? align_up((size_t)N * sizeof(unsigned long long), 256) + 256 : 0;
}

// the workspace of the step calls must be zero before its FIRST use (the kernel leaves it zeroed)
extern "C" int emo_rnnt_joint_step(const float* enc_proj, const int* enc_row, const float* dec_proj, const float* w_out,
                                   const float* b_out, int N, int J, int V, float* logits, long long* token, void* ws,
                                   size_t ws_bytes, void* stream) {
    EMO_REQUIRE(logits || token, EMO_BAD_ARG, "joint_step: neither logits nor token requested");
    StepArgs a = {};
    a.enc = enc_proj; a.enc_row = enc_row; a.dec = dec_proj; a.w = w_out; a.b = b_out;
    a.logits = logits; a.token = token; a.N = N; a.J = J; a.V = V;
    return step_launch(a, ws, ws_bytes, token != nullptr, (cudaStream_t)stream);
}

extern "C" int emo_rnnt_greedy_step(const float* enc_proj, const float* dec_proj, const float* w_out, const float* b_out,
                                    const int* tlen, int N, int T, int J, int V, int blank, int max_len, int* t_idx,
                                    int* hyp, int* hyp_len, int* align, int* align_len, int align_cap,
                                    unsigned char* emitted, long long* token, int* n_active, void* ws, size_t ws_bytes,
                                    void* stream) {
    EMO_REQUIRE(tlen && t_idx && hyp && hyp_len && emitted && token, EMO_BAD_ARG, "greedy_step: null pointer");
    EMO_REQUIRE(T > 0 && max_len > 0 && blank >= 0 && blank < V, EMO_BAD_ARG, "greedy_step: bad sizes");
    EMO_REQUIRE(!align || (align_len && align_cap > 0), EMO_BAD_ARG, "greedy_step: align needs align_len / align_cap");
    StepArgs a = {};
    a.enc = enc_proj; a.dec = dec_proj; a.w = w_out; a.b = b_out; a.token = token; a.N = N; a.J = J; a.V = V;
    a.t_idx = t_idx; a.tlen = tlen; a.T = T; a.hyp = hyp; a.hyp_len = hyp_len; a.align = align; a.align_len = align_len;
    a.emitted = emitted; a.n_active = n_active; a.blank = blank; a.max_len = max_len; a.align_cap = align_cap;
    return step_launch(a, ws, ws_bytes, true, (cudaStream_t)stream);
}

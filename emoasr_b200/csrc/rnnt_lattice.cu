// RNN-T alpha/beta lattice as an anti-diagonal wavefront + transition posteriors.
//
// Replaces the alpha/beta/grad kernels behind warp_rnnt.rnnt_loss
// (asr/modeling/decoders/rnn_transducer.py:106-115) and the Numba spin-lock kernels of
// asr/modeling/decoders/rnnt_aligner.py:14-152 (same recursion).
//
// One CTA per (utterance, direction).  Thread u owns column u of the lattice and walks the
// anti-diagonals d = t + u.  The value coming from (t-1,u) stays in a register, the value coming
// from (t,u-1) is handed over by the neighbouring thread: inside a warp with one shuffle, across
// warp boundaries through a double-buffered shared-memory slot.  The per-cell {blank,label}
// log-prob pair is one 8-byte load that does not depend on the recursion, so it is prefetched
// one block of kPrefetch diagonals ahead in registers.
#include "common.cuh"

namespace emo {

namespace {

constexpr int kPrefetch = 8;

template <bool kBackward>
__device__ __forceinline__ void wavefront(const float2* __restrict__ lp2_b, int T_b, int U_b,
                                          int U1, float* __restrict__ out_b,
                                          float* __restrict__ cost_out) {
    // shared slot per warp boundary: value handed from the last lane of warp w to lane 0 of
    // warp w+1 (forward) or from lane 0 of warp w to lane 31 of warp w-1 (backward)
    __shared__ float edge[2][33];
    const int u = threadIdx.x;
    const int lane = u & 31, warp = u >> 5;
    const int n_diag = T_b + U_b;  // diagonals 0 .. T_b+U_b-1
    const bool col_valid = u <= U_b;

    // my cell on diagonal index i (processing order): forward d=i, backward d=n_diag-1-i
    auto cell_t = [&](int i) { return (kBackward ? (n_diag - 1 - i) : i) - u; };
    // clamped address, no select on the loaded pair (it is only consumed when the cell is active):
    // a select right behind the load would stall the thread until it lands and defeat the prefetch
    const int u_ld = col_valid ? u : 0;
    auto load = [&](int i) -> float2 {
        const int t = min(max(cell_t(min(i, n_diag - 1)), 0), T_b - 1);
        return __ldg(&lp2_b[(size_t)t * U1 + u_ld]);
    };

    // The pairs of the NEXT block of kPrefetch diagonals are loaded at the top of the current block (kPrefetch
    // independent loads back to back), so every load has a whole block of diagonals to land.  (A per-diagonal
    // ring "load diagonal i + kPrefetch while computing diagonal i" was scheduled by the compiler right in front
    // of its use: a quarter of the kernel's stall samples sat on the load's scoreboard.)
    float2 cur[kPrefetch], nxt[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) cur[k] = load(k);

    float keep = kNegInf;  // forward: alpha(t-1,u)+blank(t-1,u); backward: beta(t+1,u)
    float give = kNegInf;  // forward: alpha(t,u)+label(t,u) for thread u+1; backward: beta(t,u) for u-1

    for (int i0 = 0; i0 < n_diag; i0 += kPrefetch) {
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) nxt[k] = load(i0 + kPrefetch + k);
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int i = i0 + k;
            if (i >= n_diag) break;  // uniform across the CTA
            const float2 lp = cur[k];
            const int t = cell_t(i);
            const bool active = col_valid && t >= 0 && t < T_b;

            // neighbour hand-over of the previous diagonal's `give`
            float nb = kBackward ? __shfl_down_sync(0xffffffffu, give, 1)
                                 : __shfl_up_sync(0xffffffffu, give, 1);
            if (!kBackward && lane == 0) nb = warp > 0 ? edge[(i + 1) & 1][warp - 1] : kNegInf;
            if (kBackward && lane == 31) nb = edge[(i + 1) & 1][warp + 1];

            float val = kNegInf;
            if (active) {
                if (!kBackward) {
                    // alpha(t,u) = lse(alpha(t-1,u)+blank(t-1,u), alpha(t,u-1)+label(t,u-1))
                    val = (t == 0 && u == 0) ? 0.f : log_add_exp_fast(keep, u > 0 ? nb : kNegInf);
                    out_b[(size_t)t * U1 + u] = val;
                    keep = val + lp.x;
                    give = (u < U_b) ? val + lp.y : kNegInf;
                    if (t == T_b - 1 && u == U_b) *cost_out = -(val + lp.x);
                } else {
                    // beta(t,u) = lse(beta(t+1,u)+blank(t,u), beta(t,u+1)+label(t,u))
                    if (t == T_b - 1 && u == U_b) {
                        val = lp.x;
                    } else {
                        float ne = (t < T_b - 1) ? keep + lp.x : kNegInf;
                        float em = (u < U_b) ? nb + lp.y : kNegInf;
                        val = log_add_exp_fast(ne, em);
                    }
                    out_b[(size_t)t * U1 + u] = val;
                    keep = val;
                    give = val;
                }
            } else {
                give = kNegInf;
            }
            if (!kBackward && lane == 31) edge[i & 1][warp] = give;
            if (kBackward && lane == 0) edge[i & 1][warp] = give;
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) cur[k] = nxt[k];
    }
}

// grid (B, 2): y=0 alpha, y=1 beta.  blockDim.x = U1 rounded up to a warp (<= 1024).
__global__ void __launch_bounds__(1024, 1)
rnnt_alpha_beta_kernel(const float* __restrict__ lp2, const int* __restrict__ tlen,
                       const int* __restrict__ ulen, int T, int U1, float* __restrict__ alpha,
                       float* __restrict__ beta, float* __restrict__ cost) {
    const int b = blockIdx.x;
    int T_b = min(max(tlen[b], 1), T);
    int U_b = min(max(ulen[b], 0), U1 - 1);
    const float2* lp2_b = reinterpret_cast<const float2*>(lp2) + (size_t)b * T * U1;
    float dummy;
    if (blockIdx.y == 0)
        wavefront<false>(lp2_b, T_b, U_b, U1, alpha + (size_t)b * T * U1, cost + b);
    else
        wavefront<true>(lp2_b, T_b, U_b, U1, beta + (size_t)b * T * U1, &dummy);
}

// One thread per lattice cell: posteriors of the two outgoing transitions.
__global__ void rnnt_gamma_kernel(const float* __restrict__ lp2, const int* __restrict__ tlen,
                                  const int* __restrict__ ulen, int B, int T, int U1,
                                  const float* __restrict__ alpha, const float* __restrict__ beta,
                                  const float* __restrict__ cost, float* __restrict__ gamma2) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)B * T * U1;
    if (idx >= n) return;
    int u = (int)(idx % U1);
    int t = (int)((idx / U1) % T);
    int b = (int)(idx / ((size_t)U1 * T));
    int T_b = min(max(tlen[b], 1), T);
    int U_b = min(max(ulen[b], 0), U1 - 1);
    float2 g = make_float2(0.f, 0.f);
    float c = cost[b];  // = -ll
    if (t < T_b && u <= U_b && c < INFINITY && c == c) {
        float2 lp = reinterpret_cast<const float2*>(lp2)[idx];
        float a = alpha[idx];
        float bt = (t < T_b - 1) ? beta[idx + U1] : (u == U_b ? 0.f : kNegInf);
        g.x = expf(a + lp.x + bt + c);
        if (u < U_b) g.y = expf(a + lp.y + beta[idx + 1] + c);
    }
    reinterpret_cast<float2*>(gamma2)[idx] = g;
}

// ---- dense seam (warp_rnnt-compatible) ----
__global__ void rnnt_gather_kernel(const float* __restrict__ log_probs,
                                   const int* __restrict__ labels, const int* __restrict__ tlen,
                                   const int* __restrict__ ulen, int B, int T, int U1, int V,
                                   int blank, float* __restrict__ lp2) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)B * T * U1;
    if (idx >= n) return;
    int u = (int)(idx % U1);
    int t = (int)((idx / U1) % T);
    int b = (int)(idx / ((size_t)U1 * T));
    int T_b = min(max(tlen[b], 1), T);
    int U_b = min(max(ulen[b], 0), U1 - 1);
    float2 o = make_float2(0.f, 0.f);
    if (t < T_b && u <= U_b) {
        const float* row = log_probs + idx * V;
        o.x = __ldg(row + blank);
        if (u < U_b) {
            int y = labels[(size_t)b * (U1 - 1) + u];
            y = min(max(y, 0), V - 1);
            o.y = __ldg(row + y);
        }
    }
    reinterpret_cast<float2*>(lp2)[idx] = o;
}

__global__ void rnnt_scatter_grad_kernel(const float* __restrict__ gamma2,
                                         const int* __restrict__ labels,
                                         const int* __restrict__ tlen,
                                         const int* __restrict__ ulen,
                                         const float* __restrict__ grad_cost, int B, int T, int U1,
                                         int V, int blank, float* __restrict__ grad) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)B * T * U1;
    if (idx >= n) return;
    int u = (int)(idx % U1);
    int t = (int)((idx / U1) % T);
    int b = (int)(idx / ((size_t)U1 * T));
    int T_b = min(max(tlen[b], 1), T);
    int U_b = min(max(ulen[b], 0), U1 - 1);
    if (t >= T_b || u > U_b) return;
    float2 g = reinterpret_cast<const float2*>(gamma2)[idx];
    float s = grad_cost[b];
    float* row = grad + idx * V;
    float gb = -s * g.x;
    if (u < U_b) {
        int y = labels[(size_t)b * (U1 - 1) + u];
        y = min(max(y, 0), V - 1);
        if (y == blank) gb += -s * g.y;
        else row[y] = -s * g.y;
    }
    row[blank] = gb;
}

// Forced alignment on the lattice (asr/modeling/decoders/rnnt_aligner.py:176-196): walk from (0,0), at every cell
// compare the path mass alpha + beta of the two successors -- (t+1,u) consumes a frame, (t,u+1) emits y_u at frame t.
// One thread per utterance: T_b + U_b dependent steps on L2-resident values; entries for labels the walk does not
// reach (it stops at the last frame) keep the reference's initial 0.
__global__ void rnnt_align_kernel(const float* __restrict__ alpha, const float* __restrict__ beta,
                                  const int* __restrict__ tlen, const int* __restrict__ ulen, int B, int T, int U1,
                                  int* __restrict__ aligns) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int T_b = min(max(tlen[b], 1), T), U_b = min(max(ulen[b], 0), U1 - 1);
    const float* a = alpha + (size_t)b * T * U1;
    const float* bt = beta + (size_t)b * T * U1;
    int* out = aligns + (size_t)b * (U1 - 1);
    for (int u = 0; u < U1 - 1; ++u) out[u] = 0;
    int t = 0, u = 0;
    while (t + 1 < T_b && u < U_b) {
        const size_t down = (size_t)(t + 1) * U1 + u, right = (size_t)t * U1 + u + 1;
        if (a[down] + bt[down] > a[right] + bt[right]) {
            ++t;
        } else {
            out[u] = t;
            ++u;
        }
    }
}

}  // namespace

int rnnt_lattice_launch(const float* lp2, const int* tlen, const int* ulen, int B, int T, int U1,
                        float* alpha_ws, float* beta_ws, float* cost, float* gamma2,
                        cudaStream_t st) {
    EMO_REQUIRE(lp2 && tlen && ulen && alpha_ws && beta_ws && cost && gamma2, EMO_BAD_ARG,
                "rnnt_lattice: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0, EMO_BAD_ARG, "rnnt_lattice: B,T,U1 must be positive");
    EMO_REQUIRE(U1 <= 1024, EMO_UNSUPPORTED_SHAPE,
                "rnnt_lattice: U+1 = %d exceeds 1024 lattice columns", U1);
    int threads = (U1 + 31) / 32 * 32;
    rnnt_alpha_beta_kernel<<<dim3(B, 2), threads, 0, st>>>(lp2, tlen, ulen, T, U1, alpha_ws,
                                                           beta_ws, cost);
    EMO_CHECK_LAUNCH("rnnt_alpha_beta_kernel");
    size_t n = (size_t)B * T * U1;
    rnnt_gamma_kernel<<<ceil_div(n, 256), 256, 0, st>>>(lp2, tlen, ulen, B, T, U1, alpha_ws,
                                                        beta_ws, cost, gamma2);
    EMO_CHECK_LAUNCH("rnnt_gamma_kernel");
    return EMO_OK;
}

}  // namespace emo

using namespace emo;

extern "C" int emo_rnnt_lattice_fwd_bwd(const float* lp2, const int* tlen, const int* ulen, int B,
                                        int T, int U1, float* alpha_ws, float* beta_ws,
                                        float* cost, float* gamma2, void* stream) {
    return rnnt_lattice_launch(lp2, tlen, ulen, B, T, U1, alpha_ws, beta_ws, cost, gamma2,
                               (cudaStream_t)stream);
}

extern "C" int emo_rnnt_align(const float* alpha_ws, const float* beta_ws, const int* tlen, const int* ulen, int B,
                              int T, int U1, int* aligns, void* stream) {
    EMO_REQUIRE(alpha_ws && beta_ws && tlen && ulen && (aligns || U1 == 1), EMO_BAD_ARG, "rnnt_align: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0, EMO_BAD_ARG, "rnnt_align: B,T,U1 must be positive");
    if (U1 == 1) return EMO_OK;
    rnnt_align_kernel<<<ceil_div(B, 32), 32, 0, (cudaStream_t)stream>>>(alpha_ws, beta_ws, tlen, ulen, B, T, U1, aligns);
    EMO_CHECK_LAUNCH("rnnt_align_kernel");
    return EMO_OK;
}

extern "C" int emo_rnnt_dense_fwd(const float* log_probs, const int* labels, const int* tlen,
                                  const int* ulen, int B, int T, int U1, int V, int blank,
                                  float* lp2_ws, float* alpha_ws, float* beta_ws, float* cost,
                                  float* gamma2_ws, void* stream) {
    EMO_REQUIRE(log_probs && labels && lp2_ws, EMO_BAD_ARG, "rnnt_dense_fwd: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0 && V > 0, EMO_BAD_ARG, "rnnt_dense_fwd: bad sizes");
    EMO_REQUIRE(blank >= 0 && blank < V, EMO_BAD_ARG, "rnnt_dense_fwd: blank %d outside [0,%d)",
                blank, V);
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)B * T * U1;
    rnnt_gather_kernel<<<ceil_div(n, 256), 256, 0, st>>>(log_probs, labels, tlen, ulen, B, T, U1,
                                                         V, blank, lp2_ws);
    EMO_CHECK_LAUNCH("rnnt_gather_kernel");
    return rnnt_lattice_launch(lp2_ws, tlen, ulen, B, T, U1, alpha_ws, beta_ws, cost, gamma2_ws,
                               st);
}

extern "C" int emo_rnnt_dense_bwd(const float* gamma2_ws, const int* labels, const int* tlen,
                                  const int* ulen, const float* grad_cost, int B, int T, int U1,
                                  int V, int blank, float* grad_log_probs, void* stream) {
    EMO_REQUIRE(gamma2_ws && labels && tlen && ulen && grad_cost && grad_log_probs, EMO_BAD_ARG,
                "rnnt_dense_bwd: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0 && V > 0 && blank >= 0 && blank < V, EMO_BAD_ARG,
                "rnnt_dense_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)B * T * U1;
    EMO_CUDA(cudaMemsetAsync(grad_log_probs, 0, n * V * sizeof(float), st));
    rnnt_scatter_grad_kernel<<<ceil_div(n, 256), 256, 0, st>>>(gamma2_ws, labels, tlen, ulen,
                                                               grad_cost, B, T, U1, V, blank,
                                                               grad_log_probs);
    EMO_CHECK_LAUNCH("rnnt_scatter_grad_kernel");
    return EMO_OK;
}

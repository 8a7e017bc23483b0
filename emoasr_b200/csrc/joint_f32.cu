// EMO_PREC_FP32: the joint + log-softmax of asr/modeling/decoders/rnn_transducer.py:147-156,102 and
// its backward in plain fp32 FFMA arithmetic -- the parity mode (1e-5 loss / 1e-4 grads against
// the reference).  Works for any J and V.
//
// The (B,T,U1,V) logits are streamed through a bounded slab of `slab_rows` (b,t) rows (each U1
// cells) that lives in the caller's workspace, so memory stays O(slab) instead of O(B T U V):
//   forward   h = tanh(enc+dec) -> z = h W^T + b -> row lse, gather {blank,label}
//   backward  recompute h, z -> dz in place -> d_w_out += dz^T h, d_b_out += colsum(dz),
//             dh = dz W -> dpre = dh (1-h^2) -> d_enc_proj (sum over u), d_dec_proj (sum over t)
// Slabs are processed in stream order, so all accumulations are deterministic (no atomics except
// the per-CTA column sums of d_b_out).
#include "common.cuh"

namespace emo {
namespace {

constexpr int kSlabCells = 16384;

inline int slab_rows_for(int U1) { return max(1, kSlabCells / U1); }

// ---------------- generic tiled SGEMM: C[M,N] (+)= A(M,K) B(K,N) (+ bias[N]) ----------------
// A_KC: A stored [M][K] (k contiguous) else [K][M];  B_KC: B stored [N][K] else [K][N].
template <bool A_KC, bool B_KC, bool ACCUM, bool BIAS>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
             const float* __restrict__ bias, int M, int N, int K, int lda, int ldb, int ldc) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        if (A_KC) {
            int m = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int gm = m0 + m, gk = k0 + kk + i;
                As[kk + i][m] = (gm < M && gk < K) ? __ldg(A + (size_t)gm * lda + gk) : 0.f;
            }
        } else {
            int kk = tid >> 4, m = (tid & 15) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int gm = m0 + m + i, gk = k0 + kk;
                As[kk][m + i] = (gm < M && gk < K) ? __ldg(A + (size_t)gk * lda + gm) : 0.f;
            }
        }
        if (B_KC) {
            int n = tid >> 2, kk = (tid & 3) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int gn = n0 + n, gk = k0 + kk + i;
                Bs[kk + i][n] = (gn < N && gk < K) ? __ldg(Bm + (size_t)gn * ldb + gk) : 0.f;
            }
        } else {
            int kk = tid >> 4, n = (tid & 15) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int gn = n0 + n + i, gk = k0 + kk;
                Bs[kk][n + i] = (gn < N && gk < K) ? __ldg(Bm + (size_t)gk * ldb + gn) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (BIAS) v += bias[gn];
            float* c = C + (size_t)gm * ldc + gn;
            if (ACCUM) v += *c;
            *c = v;
        }
    }
}

template <bool A_KC, bool B_KC, bool ACCUM, bool BIAS>
void sgemm(const float* A, const float* Bm, float* C, const float* bias, int M, int N, int K,
           int lda, int ldb, int ldc, cudaStream_t st) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    sgemm_kernel<A_KC, B_KC, ACCUM, BIAS><<<grid, 256, 0, st>>>(A, Bm, C, bias, M, N, K, lda, ldb, ldc);
}

// h_slab[(r-r0)*U1+u, j] = tanh(enc[r,j] + dec[b*U1+u, j]), r = b*T+t
__global__ void hidden_slab_kernel(const float* __restrict__ enc, const float* __restrict__ dec,
                                   int r0, int rows, int T, int U1, int J, float* __restrict__ h) {
    size_t n = (size_t)rows * U1 * J;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        int j = (int)(i % J);
        size_t c = i / J;
        int u = (int)(c % U1);
        int r = r0 + (int)(c / U1);
        int b = r / T;
        h[i] = tanhf(__ldg(enc + (size_t)r * J + j) + __ldg(dec + ((size_t)b * U1 + u) * J + j));
    }
}

// one warp per cell row of the slab: lse over V, gather blank/label
__global__ void __launch_bounds__(256)
row_lse_gather_kernel(const float* __restrict__ z, const int* __restrict__ labels,
                      const int* __restrict__ tlen, const int* __restrict__ ulen, int r0, int rows,
                      int T, int U1, int V, int blank, float* __restrict__ lp2,
                      float* __restrict__ lse) {
    int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int ncell = rows * U1;
    if (wid >= ncell) return;
    int u = wid % U1, r = r0 + wid / U1;
    int b = r / T, t = r - b * T;
    int T_b = min(max(tlen[b], 1), T), U_b = min(max(ulen[b], 0), U1 - 1);
    size_t cell = (size_t)r * U1 + u;
    if (t >= T_b || u > U_b) {
        if (lane == 0) {
            reinterpret_cast<float2*>(lp2)[cell] = make_float2(0.f, 0.f);
            lse[cell] = 0.f;
        }
        return;
    }
    const float* row = z + (size_t)wid * V;
    float m = kNegInf, s = 0.f;
    for (int v = lane; v < V; v += 32) {
        float x = row[v];
        float mn = fmaxf(m, x);
        if (mn > kNegInf) { s = s * expf(m - mn) + expf(x - mn); m = mn; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    if (lane == 0) {
        float l = m + logf(s);
        float2 o = make_float2(row[blank] - l, 0.f);
        if (u < U_b) {
            int y = min(max(labels[(size_t)b * (U1 - 1) + u], 0), V - 1);
            o.y = row[y] - l;
        }
        reinterpret_cast<float2*>(lp2)[cell] = o;
        lse[cell] = l;
    }
}

// dz in place + column sums.  CTA = 64 cell rows x all columns (thread per column, strided).
__global__ void __launch_bounds__(256)
dz_kernel(float* __restrict__ z, const int* __restrict__ labels, const float* __restrict__ lse,
          const float* __restrict__ gamma2, const float* __restrict__ grad_cost,
          const float* __restrict__ grad_lse, int r0, int rows,
          int T, int U1, int V, int blank, float* __restrict__ d_b_out) {
    constexpr int R = 64;
    __shared__ float s_lse[R], s_gb[R], s_gl[R], s_g[R], s_gs[R];
    __shared__ int s_lab[R];
    int ncell = rows * U1;
    int c0 = blockIdx.x * R;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        int c = c0 + i;
        if (c < ncell) {
            int u = c % U1, r = r0 + c / U1;
            int b = r / T;
            size_t cell = (size_t)r * U1 + u;
            float2 g = reinterpret_cast<const float2*>(gamma2)[cell];
            s_lse[i] = lse[cell];
            s_gb[i] = g.x;
            s_gl[i] = g.y;
            s_g[i] = grad_cost[b];
            s_gs[i] = grad_lse ? grad_lse[cell] : 0.f;   // d lse / d z = softmax(z)
            s_lab[i] = (u < U1 - 1) ? min(max(labels[(size_t)b * (U1 - 1) + u], 0), V - 1) : -1;
        }
    }
    __syncthreads();
    int nr = min(R, ncell - c0);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float colsum = 0.f;
        for (int i = 0; i < nr; ++i) {
            float* p = z + (size_t)(c0 + i) * V + v;
            float gb = s_gb[i], gl = s_gl[i], gs = s_gs[i];
            float d = 0.f;
            if (gb != 0.f || gl != 0.f || gs != 0.f) {
                const float pr = expf(*p - s_lse[i]);
                d = (gb + gl) * pr;
                if (v == blank) d -= gb;
                if (v == s_lab[i]) d -= gl;
                d = fmaf(d, s_g[i], gs * pr);
            }
            *p = d;
            colsum += d;
        }
        atomicAdd(d_b_out + v, colsum);
    }
}

// d_enc[r, j] = sum_u dh[(r-r0)*U1+u, j] * (1 - h^2);  also rewrites dh <- dpre for the next kernel
__global__ void dpre_enc_kernel(float* __restrict__ dh, const float* __restrict__ h, int r0,
                                int rows, int U1, int J, float* __restrict__ d_enc) {
    size_t n = (size_t)rows * J;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = (int)(i % J);
    int rl = (int)(i / J);
    float acc = 0.f;
    for (int u = 0; u < U1; ++u) {
        size_t k = ((size_t)rl * U1 + u) * J + j;
        float hv = h[k];
        float d = dh[k] * (1.f - hv * hv);
        dh[k] = d;
        acc += d;
    }
    d_enc[(size_t)(r0 + rl) * J + j] = acc;
}

// d_dec[b*U1+u, j] += sum over slab rows of utterance b of dpre
__global__ void dpre_dec_kernel(const float* __restrict__ dpre, int r0, int rows, int T, int U1,
                                int J, float* __restrict__ d_dec) {
    size_t n = (size_t)U1 * J;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = (int)(i % J);
    int u = (int)(i / J);
    int cur_b = r0 / T;
    float acc = 0.f;
    for (int rl = 0; rl < rows; ++rl) {
        int b = (r0 + rl) / T;
        if (b != cur_b) {
            d_dec[((size_t)cur_b * U1 + u) * J + j] += acc;
            acc = 0.f;
            cur_b = b;
        }
        acc += dpre[((size_t)rl * U1 + u) * J + j];
    }
    d_dec[((size_t)cur_b * U1 + u) * J + j] += acc;
}

int check_joint_args(const void* enc, const void* dec, const void* w, const void* bo,
                     const void* labels, const void* tlen, const void* ulen, int B, int T, int U1,
                     int J, int V, int blank) {
    EMO_REQUIRE(enc && dec && w && bo && labels && tlen && ulen, EMO_BAD_ARG, "joint: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0 && J > 0 && V > 0, EMO_BAD_ARG, "joint: bad sizes");
    EMO_REQUIRE(blank >= 0 && blank < V, EMO_BAD_ARG, "joint: blank %d outside [0,%d)", blank, V);
    return EMO_OK;
}

}  // namespace

size_t joint_f32_workspace(int op, int B, int T, int U1, int J, int V) {
    size_t cells = (size_t)min(slab_rows_for(U1), B * T) * U1;
    size_t bytes = align_up(cells * J * sizeof(float), 256) + align_up(cells * V * sizeof(float), 256);
    if (op == EMO_OP_RNNT_JOINT_BWD) bytes += align_up(cells * J * sizeof(float), 256);
    return bytes;
}

int joint_f32_launches(int op, int B, int T, int U1, int J, int V) {
    (void)J; (void)V;
    int R = min(slab_rows_for(U1), B * T);
    int slabs = (B * T + R - 1) / R;
    return slabs * (op == EMO_OP_RNNT_JOINT_BWD ? 7 : 3);
}

int joint_fwd_f32(const float* enc_proj, const float* dec_proj, const float* w_out,
                  const float* b_out, const int* labels, const int* tlen, const int* ulen, int B,
                  int T, int U1, int J, int V, int blank, float* lp2, float* lse, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
    int rc = check_joint_args(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(lp2 && lse && ws, EMO_BAD_ARG, "joint_fwd: null output/workspace");
    EMO_REQUIRE(ws_bytes >= joint_f32_workspace(EMO_OP_RNNT_JOINT_FWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_fwd(fp32): workspace %zu < %zu", ws_bytes,
                joint_f32_workspace(EMO_OP_RNNT_JOINT_FWD, B, T, U1, J, V));
    int R = min(slab_rows_for(U1), B * T);
    size_t cells = (size_t)R * U1;
    float* h = (float*)ws;
    float* z = (float*)((char*)ws + align_up(cells * J * sizeof(float), 256));
    for (int r0 = 0; r0 < B * T; r0 += R) {
        int rows = min(R, B * T - r0);
        int nc = rows * U1;
        hidden_slab_kernel<<<min(ceil_div((size_t)nc * J, 256), sm_count() * 16), 256, 0, st>>>(
            enc_proj, dec_proj, r0, rows, T, U1, J, h);
        sgemm<true, true, false, true>(h, w_out, z, b_out, nc, V, J, J, J, V, st);
        row_lse_gather_kernel<<<ceil_div((size_t)nc * 32, 256), 256, 0, st>>>(
            z, labels, tlen, ulen, r0, rows, T, U1, V, blank, lp2, lse);
    }
    EMO_CHECK_LAUNCH("joint_fwd_f32");
    return EMO_OK;
}

int joint_bwd_f32(const float* enc_proj, const float* dec_proj, const float* w_out,
                  const float* b_out, const int* labels, const int* tlen, const int* ulen,
                  const float* lse, const float* gamma2, const float* grad_cost, const float* grad_lse, int B,
                  int T, int U1, int J, int V, int blank, float* d_enc_proj, float* d_dec_proj,
                  float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_joint_args(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(lse && gamma2 && grad_cost && d_enc_proj && d_dec_proj && d_w_out && d_b_out && ws,
                EMO_BAD_ARG, "joint_bwd: null pointer");
    EMO_REQUIRE(ws_bytes >= joint_f32_workspace(EMO_OP_RNNT_JOINT_BWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_bwd(fp32): workspace %zu < %zu", ws_bytes,
                joint_f32_workspace(EMO_OP_RNNT_JOINT_BWD, B, T, U1, J, V));
    int R = min(slab_rows_for(U1), B * T);
    size_t cells = (size_t)R * U1;
    float* h = (float*)ws;
    float* z = (float*)((char*)ws + align_up(cells * J * sizeof(float), 256));
    float* dh = (float*)((char*)z + align_up(cells * V * sizeof(float), 256));
    EMO_CUDA(cudaMemsetAsync(d_dec_proj, 0, (size_t)B * U1 * J * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_w_out, 0, (size_t)V * J * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_out, 0, (size_t)V * sizeof(float), st));
    for (int r0 = 0; r0 < B * T; r0 += R) {
        int rows = min(R, B * T - r0);
        int nc = rows * U1;
        hidden_slab_kernel<<<min(ceil_div((size_t)nc * J, 256), sm_count() * 16), 256, 0, st>>>(
            enc_proj, dec_proj, r0, rows, T, U1, J, h);
        sgemm<true, true, false, true>(h, w_out, z, b_out, nc, V, J, J, J, V, st);
        dz_kernel<<<ceil_div(nc, 64), 256, 0, st>>>(z, labels, lse, gamma2, grad_cost, grad_lse, r0, rows, T,
                                                    U1, V, blank, d_b_out);
        // d_w_out[v,j] += sum_c dz[c,v] h[c,j]
        sgemm<false, false, true, false>(z, h, d_w_out, nullptr, V, J, nc, V, J, J, st);
        // dh[c,j] = sum_v dz[c,v] w_out[v,j]
        sgemm<true, false, false, false>(z, w_out, dh, nullptr, nc, J, V, V, J, J, st);
        dpre_enc_kernel<<<ceil_div((size_t)rows * J, 256), 256, 0, st>>>(dh, h, r0, rows, U1, J,
                                                                        d_enc_proj);
        dpre_dec_kernel<<<ceil_div((size_t)U1 * J, 256), 256, 0, st>>>(dh, r0, rows, T, U1, J,
                                                                      d_dec_proj);
    }
    EMO_CHECK_LAUNCH("joint_bwd_f32");
    return EMO_OK;
}

}  // namespace emo

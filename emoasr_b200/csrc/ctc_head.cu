// Fused CTC head: output Linear(He,V) + log_softmax + CTC loss, forward and backward, with the
// (B,T,V) logits / log-probs / gradient never written to memory (asr/modeling/decoders/ctc.py:34,103-113).
//
// The head is the transducer joint with one "cell" per frame and no tanh: z[b,t,:] = W e[b,t] + b.  The dense,
// vocabulary-sized work therefore runs on the joint's tensor-core kernels in their `plain` mode (h = enc stream):
//   forward   joint_fwd_kernel<plain>   z tiles in TMEM, online log-sum-exp -> lse[b,t] and the blank log-prob
//   backward  joint_bwd_ring_kernel     recomputes z tiles, dz = g softmax(z) handed through the L2-resident ring to
//                                       the d_eouts (dz W) and d_W (dz^T e) GEMMs; d_b = column sums
// What CTC adds is SPARSE: per frame only the <= U_b + 1 distinct entries of the blank-extended label sequence carry
// an emission (forward) or a state posterior (backward).  Those are small per-utterance contractions against the
// gathered weight rows W[y_u], done by the CUDA-core kernels of this file:
//   head_emission_kernel   emis[b,t,1+u] = bf16(e[b,t]) . bf16(W[y_u]) + b[y_u] - lse[b,t]   (same operand rounding as the
//                          tensor-core logits, so emission and lse are consistent); staged as the lattice's input
//   head_deouts_kernel     occ = exp(alpha + beta - emis - ll) per (frame, label) -> g occ kept for the dW pass;
//                          d_eouts[b,t,:] = dh[b,t,:] - g sum_j occ_j W[row_j,:];  d_b[row_j] -= g sum_t occ_j
//   head_dw_kernel         d_W[row_j,:] -= sum_t (g occ_j)[t] e[b,t,:]
// with row_0 = blank (all blank states of a frame summed) and row_{1+u} = y_u.  The alpha / beta recursions are the
// kernels of ctc.cu (ctc_lattice_launch).
#include "joint_tc.cuh"

namespace emo {
namespace {

constexpr int kHeadFrames = 8;      // frames per block of the emission / d_eouts kernels
constexpr int kHeadThreads = 128;

// The joint kernels tile the cells of one utterance at a time (256 per CTA pair).  With one cell per frame that
// would waste the tail of every utterance's last tile (T = 374 -> 512 rows), so the head presents the batch as Bs
// "super-utterances" of Ts = (B / Bs) * T frames each (Bs = 1 whenever B * T < 65536): padded frames are ordinary
// rows whose per-row gradient scale is 0.
struct Super {
    int Bs, Ts;
};
Super head_super(int B, int T) {
    Super s;
    s.Bs = B;
    for (int d = 1; d <= B; ++d)
        if (B % d == 0 && (long long)(B / d) * T < 65536) { s.Bs = d; break; }
    s.Ts = B / s.Bs * T;
    return s;
}

struct HeadWs {
    __nv_bfloat16* w_bf16;   // (Vp, He)
    __half* e16;             // (B, T, He)
    float* b_pad;            // (Vp)
    int* tlen32;             // (B)  clamp(tlen, 1, T)
    int* ulen32;             // (B)  0: one cell per frame
    int* tsup32;             // (Bs) Ts: frames of a super-utterance
    float* lp2;              // fwd: (B, T, 2) blank log-prob from the joint forward
    float* geff;             // bwd: (B) grad_nll, 0 for infeasible utterances
    float* grow;             // bwd: (B, T) per-frame gradient scale: geff[b] for t < T_b, else 0
    float* ll;               // bwd: (B) log-likelihood from the last alphas
    float* occg;             // bwd: (B, T, Umax + 1) g * occupancy per (frame, {blank, label u})
    void* dh;                // bwd: tile-major bf16 dh of the ring kernel
    void* ring;              // bwd: dz / h ring + flags
    size_t total;
};

HeadWs head_ws_layout(void* base, int op, int B, int T, int He, int V, int Umax) {
    HeadWs w;
    const size_t Vp = (size_t)padded_vocab(V);
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 256); return r; };
    w.w_bf16 = reinterpret_cast<__nv_bfloat16*>(take(Vp * He * sizeof(__nv_bfloat16)));
    w.e16 = reinterpret_cast<__half*>(take((size_t)B * T * He * sizeof(__half)));
    w.b_pad = reinterpret_cast<float*>(take(Vp * sizeof(float)));
    w.tlen32 = reinterpret_cast<int*>(take((size_t)B * sizeof(int)));
    w.ulen32 = reinterpret_cast<int*>(take((size_t)B * sizeof(int)));
    w.tsup32 = reinterpret_cast<int*>(take((size_t)B * sizeof(int)));
    const Super su = head_super(B, T);
    w.grow = nullptr;
    w.lp2 = nullptr; w.geff = nullptr; w.ll = nullptr; w.occg = nullptr; w.dh = nullptr; w.ring = nullptr;
    if (op == 0) {
        w.lp2 = reinterpret_cast<float*>(take((size_t)B * T * 2 * sizeof(float)));
    } else {
        w.geff = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
        w.ll = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
        w.grow = reinterpret_cast<float*>(take((size_t)B * T * sizeof(float)));
        w.occg = reinterpret_cast<float*>(take((size_t)B * T * (Umax + 1) * sizeof(float)));
        w.dh = take(align_up(dh_bytes_for(su.Bs, su.Ts, 1, He), 1024));
        w.ring = take(joint_ring_workspace(su.Bs, su.Ts, 1, He, V));
    }
    w.total = off;
    return w;
}

__device__ __forceinline__ int clamp_label(long long l, int V) { return (int)(l < 0 ? 0 : (l >= V ? V - 1 : l)); }

// lengths for the joint kernels + (backward) feasibility and the per-utterance gradient scale
__global__ void head_prep_kernel(const long long* __restrict__ tlen, const long long* __restrict__ ulen, int B, int T,
                                 int Umax, int Ts, int* __restrict__ tlen32, int* __restrict__ ulen32,
                                 int* __restrict__ tsup32, const float* __restrict__ alpha_ws,
                                 const float* __restrict__ grad_nll, float* __restrict__ geff, float* __restrict__ ll) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long T_bl = tlen[b], U_bl = ulen[b];
    const int T_b = (int)(T_bl < 1 ? 1 : (T_bl > T ? T : T_bl));
    tlen32[b] = T_b;
    ulen32[b] = 0;
    tsup32[b] = Ts;
    if (!alpha_ws) return;
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S = 2 * Umax + 1, S_b = 2 * U_b + 1;
    const float* alast = alpha_ws + ((size_t)b * T + (T_b - 1)) * S;
    const float l = log_add_exp(alast[S_b - 1], S_b > 1 ? alast[S_b - 2] : kNegInf);
    const bool feasible = l > kNegInf && l == l && l < INFINITY;
    geff[b] = feasible ? grad_nll[b] : 0.f;   // infeasible utterances: zero gradient (zero_infinity, ctc.py:38)
    ll[b] = feasible ? l : 0.f;
}

// per-frame gradient scale of the dense part
__global__ void head_grow_kernel(const int* __restrict__ tlen32, const float* __restrict__ geff, int B, int T,
                                 float* __restrict__ grow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int b = i / T, t = i - b * T;
    grow[i] = t < tlen32[b] ? geff[b] : 0.f;
}

// e as the tensor cores see it: fp32 -> fp16 (the stream cast) -> bf16 (the A producers)
__device__ __forceinline__ float round_like_a_operand(__half x) { return __bfloat162float(__float2bfloat16_rn(__half2float(x))); }

// emissions of the labels: block = (tile of kHeadFrames frames, utterance); thread = label u (stride blockDim)
__global__ void __launch_bounds__(kHeadThreads)
head_emission_kernel(const __half* __restrict__ e16, const __nv_bfloat16* __restrict__ w_bf16,
                     const float* __restrict__ b_out, const float* __restrict__ lse, const float* __restrict__ lp2,
                     const long long* __restrict__ labels, const int* __restrict__ tlen32,
                     const long long* __restrict__ ulen, int T, int He, int V, int Umax, int blank,
                     float* __restrict__ emis, float* __restrict__ lp_a, float* __restrict__ lp_b) {
    extern __shared__ float s_e[];   // [kHeadFrames][He]
    const int b = blockIdx.y, t0 = blockIdx.x * kHeadFrames;
    const int T_b = tlen32[b];
    if (t0 >= T_b) return;
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S = 2 * Umax + 1;
    const int nf = min(kHeadFrames, T_b - t0);
    for (int i = threadIdx.x; i < kHeadFrames * He; i += blockDim.x) {
        const int f = i / He, k = i - f * He;
        s_e[i] = f < nf ? round_like_a_operand(e16[((size_t)b * T + t0 + f) * He + k]) : 0.f;
    }
    __syncthreads();
    const long long* y = labels + (size_t)b * Umax;
    // blank states: the log-prob comes from the joint forward
    for (int i = threadIdx.x; i < nf * (U_b + 1); i += blockDim.x) {
        const int f = i / (U_b + 1), u = i - f * (U_b + 1);
        const size_t r = (size_t)b * T + t0 + f;
        const float v = lp2[2 * r];
        if (u == 0) emis[r * (Umax + 1)] = v;
        lp_a[r * S + 2 * u] = v;
        if (lp_b) lp_b[r * S + 2 * u] = v;
    }
    for (int u = threadIdx.x; u < U_b; u += blockDim.x) {
        const int lab = clamp_label(y[u], V);
        const uint4* wrow = reinterpret_cast<const uint4*>(w_bf16 + (size_t)lab * He);
        float acc[kHeadFrames];
#pragma unroll
        for (int f = 0; f < kHeadFrames; ++f) acc[f] = 0.f;
        for (int k8 = 0; k8 < He / 8; ++k8) {
            const uint4 wv = __ldg(wrow + k8);
            const float w0 = __uint_as_float(wv.x << 16), w1 = __uint_as_float(wv.x & 0xffff0000u);
            const float w2 = __uint_as_float(wv.y << 16), w3 = __uint_as_float(wv.y & 0xffff0000u);
            const float w4 = __uint_as_float(wv.z << 16), w5 = __uint_as_float(wv.z & 0xffff0000u);
            const float w6 = __uint_as_float(wv.w << 16), w7 = __uint_as_float(wv.w & 0xffff0000u);
#pragma unroll
            for (int f = 0; f < kHeadFrames; ++f) {
                const float4 ea = *reinterpret_cast<const float4*>(s_e + f * He + k8 * 8);
                const float4 eb = *reinterpret_cast<const float4*>(s_e + f * He + k8 * 8 + 4);
                acc[f] = fmaf(ea.x, w0, fmaf(ea.y, w1, fmaf(ea.z, w2, fmaf(ea.w, w3, acc[f]))));
                acc[f] = fmaf(eb.x, w4, fmaf(eb.y, w5, fmaf(eb.z, w6, fmaf(eb.w, w7, acc[f]))));
            }
        }
        const float bias = __ldg(b_out + lab);
#pragma unroll
        for (int f = 0; f < kHeadFrames; ++f) {
            if (f >= nf) break;
            const size_t r = (size_t)b * T + t0 + f;
            const float v = acc[f] + bias - lse[r];
            emis[r * (Umax + 1) + 1 + u] = v;
            lp_a[r * S + 2 * u + 1] = v;
            if (lp_b) lp_b[r * S + 2 * u + 1] = v;
        }
    }
}

// d_eouts = dense part (dh, bf16, from the ring kernel) - g sum_j occ_j W[row_j]; g occ kept for the dW pass;
// d_b[row_j] -= g sum_t occ_j.  Block = (tile of kHeadFrames frames, utterance); thread = 4 columns of He.
__global__ void __launch_bounds__(kHeadThreads)
head_deouts_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ w, const float* __restrict__ emis,
                   const float* __restrict__ alpha_ws, const float* __restrict__ beta_ws,
                   const long long* __restrict__ labels, const int* __restrict__ tlen32,
                   const long long* __restrict__ ulen, const float* __restrict__ geff, const float* __restrict__ ll,
                   int T, int He, int V, int Umax, int blank, int tpu, int per_super, float* __restrict__ occg,
                   float* __restrict__ d_eouts, float* __restrict__ d_b) {
    extern __shared__ float s_occ[];   // [Umax + 1][kHeadFrames]
    const int b = blockIdx.y, t0 = blockIdx.x * kHeadFrames;
    const int T_b = tlen32[b];
    const int c4 = threadIdx.x * 4;
    if (t0 >= T_b) {   // padded frames: zero gradient
        for (int f = 0; f < kHeadFrames && t0 + f < T; ++f)
            if (c4 < He) *reinterpret_cast<float4*>(d_eouts + ((size_t)b * T + t0 + f) * He + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S = 2 * Umax + 1, S_b = 2 * U_b + 1;
    const int nf = min(kHeadFrames, T_b - t0);
    const float g = geff[b], l = ll[b];
    const long long* y = labels + (size_t)b * Umax;
    for (int i = threadIdx.x; i < (Umax + 1) * kHeadFrames; i += blockDim.x) s_occ[i] = 0.f;
    __syncthreads();
    if (g != 0.f) {
        for (int i = threadIdx.x; i < S_b * kHeadFrames; i += blockDim.x) {
            const int s = i / kHeadFrames, f = i - s * kHeadFrames;
            if (f >= nf) continue;
            const size_t r = (size_t)b * T + t0 + f;
            const float a = alpha_ws[r * S + s], bt = beta_ws[r * S + s];
            if (!(a > kNegInf && bt > kNegInf)) continue;
            const int j = (s & 1) ? 1 + (s >> 1) : 0;
            const float o = g * expf(a + bt - emis[r * (Umax + 1) + j] - l);
            if (j) s_occ[j * kHeadFrames + f] = o;
            else atomicAdd(&s_occ[f], o);
        }
    }
    __syncthreads();
    // keep g occ for the dW pass; bias gradient
    for (int i = threadIdx.x; i < (U_b + 1) * kHeadFrames; i += blockDim.x) {
        const int j = i / kHeadFrames, f = i - j * kHeadFrames;
        if (f < nf) occg[((size_t)b * T + t0 + f) * (Umax + 1) + j] = s_occ[i];
    }
    if (g != 0.f) {
        for (int j = threadIdx.x; j <= U_b; j += blockDim.x) {
            float sum = 0.f;
#pragma unroll
            for (int f = 0; f < kHeadFrames; ++f) sum += s_occ[j * kHeadFrames + f];
            if (sum != 0.f) atomicAdd(d_b + (j == 0 ? blank : clamp_label(y[j - 1], V)), -sum);
        }
    }
    if (c4 >= He) return;
    float acc[kHeadFrames][4];
#pragma unroll
    for (int f = 0; f < kHeadFrames; ++f) acc[f][0] = acc[f][1] = acc[f][2] = acc[f][3] = 0.f;
    if (g != 0.f) {
        for (int j = 0; j <= U_b; ++j) {
            const int row = j == 0 ? blank : clamp_label(y[j - 1], V);
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)row * He + c4));
            const float4 oa = *reinterpret_cast<const float4*>(s_occ + j * kHeadFrames);
            const float4 ob = *reinterpret_cast<const float4*>(s_occ + j * kHeadFrames + 4);
            const float o[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
#pragma unroll
            for (int f = 0; f < kHeadFrames; ++f) {
                acc[f][0] = fmaf(o[f], wv.x, acc[f][0]);
                acc[f][1] = fmaf(o[f], wv.y, acc[f][1]);
                acc[f][2] = fmaf(o[f], wv.z, acc[f][2]);
                acc[f][3] = fmaf(o[f], wv.w, acc[f][3]);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < kHeadFrames; ++f) {
        if (t0 + f >= T) break;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < nf) {
            // tile-major dh with one cell per frame: row = super-utterance * tpu * 128 + frame inside it
            const size_t row = (size_t)(b / per_super) * tpu * kTileM + (size_t)(b % per_super) * T + t0 + f;
            const uint2 dv = *reinterpret_cast<const uint2*>(dh + row * He + c4);
            out.x = __uint_as_float(dv.x << 16) - acc[f][0];
            out.y = __uint_as_float(dv.x & 0xffff0000u) - acc[f][1];
            out.z = __uint_as_float(dv.y << 16) - acc[f][2];
            out.w = __uint_as_float(dv.y & 0xffff0000u) - acc[f][3];
        }
        *reinterpret_cast<float4*>(d_eouts + ((size_t)b * T + t0 + f) * He + c4) = out;
    }
}

// d_W[row_j,:] -= sum_t (g occ_j)[t] e[b,t,:].  Block = (8 columns j of the utterance's {blank, labels}, utterance);
// thread = 4 columns of He.
constexpr int kHeadCols = 8;
__global__ void __launch_bounds__(kHeadThreads)
head_dw_kernel(const float* __restrict__ eouts, const float* __restrict__ occg, const long long* __restrict__ labels,
               const int* __restrict__ tlen32, const long long* __restrict__ ulen, const float* __restrict__ geff,
               int T, int He, int V, int Umax, int blank, float* __restrict__ d_w) {
    const int b = blockIdx.y, j0 = blockIdx.x * kHeadCols;
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    if (j0 > U_b || geff[b] == 0.f) return;
    const int c4 = threadIdx.x * 4;
    if (c4 >= He) return;
    const int T_b = tlen32[b];
    const int nj = min(kHeadCols, U_b + 1 - j0);
    float acc[kHeadCols][4];
#pragma unroll
    for (int j = 0; j < kHeadCols; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const float* ob = occg + (size_t)b * T * (Umax + 1) + j0;
    const float* eb = eouts + (size_t)b * T * He + c4;
#pragma unroll 4
    for (int t = 0; t < T_b; ++t) {
        const float4 ev = __ldg(reinterpret_cast<const float4*>(eb + (size_t)t * He));
        const float* o = ob + (size_t)t * (Umax + 1);
#pragma unroll
        for (int j = 0; j < kHeadCols; ++j) {
            const float ov = j < nj ? __ldg(o + j) : 0.f;
            acc[j][0] = fmaf(ov, ev.x, acc[j][0]);
            acc[j][1] = fmaf(ov, ev.y, acc[j][1]);
            acc[j][2] = fmaf(ov, ev.z, acc[j][2]);
            acc[j][3] = fmaf(ov, ev.w, acc[j][3]);
        }
    }
    const long long* y = labels + (size_t)b * Umax;
#pragma unroll
    for (int j = 0; j < kHeadCols; ++j) {
        if (j >= nj) break;
        const int jj = j0 + j;
        const int row = jj == 0 ? blank : clamp_label(y[jj - 1], V);
        float* dst = d_w + (size_t)row * He + c4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(-acc[j][0]), "f"(-acc[j][1]),
                     "f"(-acc[j][2]), "f"(-acc[j][3])
                     : "memory");
    }
}

int head_check(const void* eouts, const void* w, const void* b, const void* labels, const void* tlen, const void* ulen,
               int B, int T, int He, int V, int Umax, int blank) {
    EMO_REQUIRE(eouts && w && b && labels && tlen && ulen, EMO_BAD_ARG, "ctc_head: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && He > 0 && V > 0 && Umax >= 1, EMO_BAD_ARG, "ctc_head: bad sizes");
    EMO_REQUIRE(blank >= 0 && blank < V, EMO_BAD_ARG, "ctc_head: blank %d outside [0,%d)", blank, V);
    EMO_REQUIRE(2 * Umax + 1 <= 1024, EMO_UNSUPPORTED_SHAPE, "ctc_head: 2*Umax+1 = %d exceeds 1024 extended states",
                2 * Umax + 1);
    EMO_REQUIRE(He % 128 == 0 && He <= kMaxKBlocks * kBlockK, EMO_UNSUPPORTED_SHAPE,
                "ctc_head: enc_hidden_size %d must be a multiple of 128 and <= 512 (use the unfused CTC loss)", He);
    EMO_REQUIRE(joint_ring_supported(head_super(B, T).Bs, head_super(B, T).Ts, 1, He, V) && He / 4 <= kHeadThreads,
                EMO_UNSUPPORTED_SHAPE, "ctc_head: unsupported shape (B <= 1024, T < 65536)");
    EMO_REQUIRE((long long)B * T * He < (1ll << 31), EMO_UNSUPPORTED_SHAPE, "ctc_head: eouts exceeds 2^31 elements");
    EMO_REQUIRE(((uintptr_t)eouts & 15) == 0 && ((uintptr_t)w & 15) == 0, EMO_BAD_ARG,
                "ctc_head: eouts / weight must be 16-byte aligned");
    return EMO_OK;
}

int head_casts(const float* eouts, const float* w, const float* b, int B, int T, int He, int V, const HeadWs& ws,
               cudaStream_t st) {
    const int Vp = padded_vocab(V);
    const size_t nw = (size_t)V * He, ne = (size_t)B * T * He;
    f32_to_bf16_kernel<<<ceil_div(nw, 4 * 256), 256, 0, st>>>(w, ws.w_bf16, nw);
    EMO_CHECK_LAUNCH("f32_to_bf16_kernel");
    f32_to_f16_kernel<<<ceil_div(ne, 4 * 256), 256, 0, st>>>(eouts, ws.e16, ne);
    EMO_CHECK_LAUNCH("f32_to_f16_kernel");
    if (Vp != V) {
        const size_t n_tail = (size_t)(Vp - V) * He;
        pad_vocab_kernel<<<ceil_div(max(n_tail, (size_t)Vp), 256), 256, 0, st>>>(ws.w_bf16 + nw, n_tail, b, ws.b_pad, V, Vp);
        EMO_CHECK_LAUNCH("pad_vocab_kernel");
    }
    return EMO_OK;
}

}  // namespace
}  // namespace emo

using namespace emo;

extern "C" int emo_ctc_head_supported(int B, int T, int He, int V, int Umax) {
    if (B <= 0 || T <= 0 || He <= 0 || V <= 0 || Umax < 1) return 0;
    if (2 * Umax + 1 > 1024 || He % 128 != 0 || He > kMaxKBlocks * kBlockK) return 0;
    if ((long long)B * T * He >= (1ll << 31)) return 0;
    if (T >= 65536 || !joint_ring_supported(head_super(B, T).Bs, head_super(B, T).Ts, 1, He, V)) return 0;
    // the ring kernel needs a resident CTA pair for every 256-row slab of the vocabulary plus producers / dh pairs
    return ceil_div(padded_vocab(V), 256) + 8 <= sm_count() / 2 ? 1 : 0;
}

extern "C" size_t emo_ctc_head_workspace_bytes(int op, int B, int T, int He, int V, int Umax) {
    if (!emo_ctc_head_supported(B, T, He, V, Umax) || (op != 0 && op != 1)) return 0;
    return head_ws_layout(nullptr, op, B, T, He, V, Umax).total;
}

extern "C" int emo_ctc_head_fwd(const float* eouts, const float* w, const float* b, const long long* labels,
                                const long long* tlen, const long long* ulen, int B, int T, int He, int V, int Umax,
                                int blank, int zero_infinity, float* lse, float* emis, float* alpha_ws,
                                float* beta_ws, float* nll, void* ws, size_t ws_bytes, void* stream) {
    int rc = head_check(eouts, w, b, labels, tlen, ulen, B, T, He, V, Umax, blank);
    if (rc) return rc;
    EMO_REQUIRE(lse && emis && alpha_ws && nll && ws, EMO_BAD_ARG, "ctc_head_fwd: null output pointer");
    const HeadWs L = head_ws_layout(ws, 0, B, T, He, V, Umax);
    EMO_REQUIRE(ws_bytes >= L.total && ((uintptr_t)ws & 255) == 0, EMO_WORKSPACE_TOO_SMALL,
                "ctc_head_fwd: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Super su = head_super(B, T);
    head_prep_kernel<<<ceil_div(B, 128), 128, 0, st>>>(tlen, ulen, B, T, Umax, su.Ts, L.tlen32, L.ulen32, L.tsup32,
                                                       nullptr, nullptr, nullptr, nullptr);
    EMO_CHECK_LAUNCH("head_prep_kernel");
    rc = head_casts(eouts, w, b, B, T, He, V, L, st);
    if (rc) return rc;
    const int Vp = padded_vocab(V);
    const float* bias = Vp != V ? L.b_pad : b;
    // dense part: lse[b,t] and the blank log-prob (lp2[..., 0]); labels are unused with one cell per frame
    // (padded frames are ordinary rows here: their lse / blank log-prob are computed and never read)
    rc = joint_fwd_launch(L.w_bf16, L.e16, nullptr, bias, L.ulen32, L.tsup32, L.ulen32, su.Bs, su.Ts, 1, He, Vp, blank,
                          L.lp2, lse, 1, st);
    if (rc) return rc;
    const size_t smem = (size_t)kHeadFrames * He * sizeof(float);
    head_emission_kernel<<<dim3(ceil_div(T, kHeadFrames), B), kHeadThreads, smem, st>>>(
        L.e16, L.w_bf16, b, lse, L.lp2, labels, L.tlen32, ulen, T, He, V, Umax, blank, emis, alpha_ws, beta_ws);
    EMO_CHECK_LAUNCH("head_emission_kernel");
    return ctc_lattice_launch(labels, tlen, ulen, B, T, V, Umax, blank, zero_infinity, alpha_ws, beta_ws, nll, st);
}

extern "C" int emo_ctc_head_bwd(const float* eouts, const float* w, const float* b, const long long* labels,
                                const long long* tlen, const long long* ulen, const float* lse, const float* emis,
                                const float* alpha_ws, const float* beta_ws, const float* grad_nll, int B, int T,
                                int He, int V, int Umax, int blank, float* d_eouts, float* d_w, float* d_b, void* ws,
                                size_t ws_bytes, void* stream) {
    int rc = head_check(eouts, w, b, labels, tlen, ulen, B, T, He, V, Umax, blank);
    if (rc) return rc;
    EMO_REQUIRE(lse && emis && alpha_ws && beta_ws && grad_nll && d_eouts && d_w && d_b && ws, EMO_BAD_ARG,
                "ctc_head_bwd: null pointer (the forward must have been given beta_ws)");
    EMO_REQUIRE(((uintptr_t)d_eouts & 15) == 0 && ((uintptr_t)d_w & 15) == 0, EMO_BAD_ARG,
                "ctc_head_bwd: d_eouts / d_w must be 16-byte aligned");
    const HeadWs L = head_ws_layout(ws, 1, B, T, He, V, Umax);
    EMO_REQUIRE(ws_bytes >= L.total && ((uintptr_t)ws & 255) == 0, EMO_WORKSPACE_TOO_SMALL,
                "ctc_head_bwd: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Super su = head_super(B, T);
    head_prep_kernel<<<ceil_div(B, 128), 128, 0, st>>>(tlen, ulen, B, T, Umax, su.Ts, L.tlen32, L.ulen32, L.tsup32,
                                                       alpha_ws, grad_nll, L.geff, L.ll);
    EMO_CHECK_LAUNCH("head_prep_kernel");
    head_grow_kernel<<<ceil_div((size_t)B * T, 256), 256, 0, st>>>(L.tlen32, L.geff, B, T, L.grow);
    EMO_CHECK_LAUNCH("head_grow_kernel");
    rc = head_casts(eouts, w, b, B, T, He, V, L, st);
    if (rc) return rc;
    const int Vp = padded_vocab(V);
    const float* bias = Vp != V ? L.b_pad : b;
    EMO_CUDA(cudaMemsetAsync(d_w, 0, (size_t)V * He * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b, 0, (size_t)V * sizeof(float), st));
    // dense part: dz = g softmax(z) -> d_W, d_b, dh (bf16, tile-major)
    rc = joint_bwd_ring_launch(L.w_bf16, L.e16, nullptr, bias, L.ulen32, L.tsup32, L.ulen32, lse, nullptr, nullptr,
                               L.grow, su.Bs, su.Ts, 1, He, Vp, V, blank, 1, L.dh, L.ring, d_w, d_b, st);
    if (rc) return rc;
    // sparse part: the entries of the blank-extended label sequence
    const size_t smem = (size_t)(Umax + 1) * kHeadFrames * sizeof(float);
    head_deouts_kernel<<<dim3(ceil_div(T, kHeadFrames), B), kHeadThreads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(L.dh), w, emis, alpha_ws, beta_ws, labels, L.tlen32, ulen, L.geff, L.ll, T,
        He, V, Umax, blank, tiles128_per_utt(su.Ts, 1), B / su.Bs, L.occg, d_eouts, d_b);
    EMO_CHECK_LAUNCH("head_deouts_kernel");
    head_dw_kernel<<<dim3(ceil_div(Umax + 1, kHeadCols), B), kHeadThreads, 0, st>>>(
        eouts, L.occg, labels, L.tlen32, ulen, L.geff, T, He, V, Umax, blank, d_w);
    EMO_CHECK_LAUNCH("head_dw_kernel");
    return EMO_OK;
}

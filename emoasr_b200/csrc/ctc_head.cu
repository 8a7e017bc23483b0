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
#include <mma.h>

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
    __nv_bfloat16* e_bf;     // (B * T + 16, He)  bf16(f16(e)): the A operand as the MMAs see it; 16 zero tail rows
    __nv_bfloat16* wy;       // (B, Up, He) gathered weight rows {blank, y_0 ..}, Up = Umax + 1 rounded up to 16
    float* bias_y;           // (B, Up)
    float* b_pad;            // (Vp)
    int* tlen32;             // (B)  clamp(tlen, 1, T)
    int* ulen32;             // (B)  0: one cell per frame
    int* tsup32;             // (Bs) Ts: frames of a super-utterance
    float* lp2;              // fwd: (B, T, 2) blank log-prob from the joint forward
    float* geff;             // bwd: (B) grad_nll, 0 for infeasible utterances
    float* grow;             // bwd: (B, T) per-frame gradient scale: geff[b] for t < T_b, else 0
    float* ll;               // bwd: (B) log-likelihood from the last alphas
    __nv_bfloat16* occg;     // bwd: (B, Tp, Up) g * occupancy per (frame, {blank, label u}), Tp = T rounded up to 16
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
    const size_t Up = (size_t)(Umax + 1 + 15) / 16 * 16, Tp = (size_t)(T + 15) / 16 * 16;
    w.e_bf = reinterpret_cast<__nv_bfloat16*>(take(((size_t)B * T + 16) * He * sizeof(__nv_bfloat16)));
    w.wy = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * Up * He * sizeof(__nv_bfloat16)));
    w.bias_y = reinterpret_cast<float*>(take((size_t)B * Up * sizeof(float)));
    const Super su = head_super(B, T);
    w.grow = nullptr;
    w.lp2 = nullptr; w.geff = nullptr; w.ll = nullptr; w.occg = nullptr; w.dh = nullptr; w.ring = nullptr;
    if (op == 0) {
        w.lp2 = reinterpret_cast<float*>(take((size_t)B * T * 2 * sizeof(float)));
    } else {
        w.geff = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
        w.ll = reinterpret_cast<float*>(take((size_t)B * sizeof(float)));
        w.grow = reinterpret_cast<float*>(take((size_t)B * T * sizeof(float)));
        w.occg = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * Tp * Up * sizeof(__nv_bfloat16)));
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

// ---- operands of the sparse (label) GEMMs ------------------------------------------------------------------------
// e16 (fp16, the joint kernels' enc stream) and e_bf = bf16(f16(e)) -- the value the tensor-core A producers feed
// the MMAs -- in one pass; padded frames (t >= T_b) are written as ZEROS in both, so that their rows contribute
// exactly nothing (0 * h, also when the caller's padding holds Inf / NaN).
__global__ void head_cast_e_kernel(const float* __restrict__ eouts, const int* __restrict__ tlen32, int T, int He,
                                   size_t n, __half* __restrict__ e16, __nv_bfloat16* __restrict__ e_bf) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const size_t row = i / He;
    const int b = (int)(row / T), t = (int)(row - (size_t)b * T);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < tlen32[b]) v = *reinterpret_cast<const float4*>(eouts + i);
    const uint2 h = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(e16 + i) = h;
    const float2 a = unpack_f16x2(h.x), c = unpack_f16x2(h.y);
    *reinterpret_cast<uint2*>(e_bf + i) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(c.x, c.y));
}

// wy[b][j][:] = w_bf16[row_j] (row_0 = blank, row_{1+u} = y_u, zero rows for j > U_b); bias_y[b][j] = b[row_j]
__global__ void head_gather_kernel(const __nv_bfloat16* __restrict__ w_bf16, const float* __restrict__ b_out,
                                   const long long* __restrict__ labels, const long long* __restrict__ ulen, int He,
                                   int V, int Umax, int Up, int blank, __nv_bfloat16* __restrict__ wy,
                                   float* __restrict__ bias_y) {
    const int b = blockIdx.y, j = blockIdx.x;
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const bool live = j <= U_b;
    const int row = !live ? 0 : (j == 0 ? blank : clamp_label(labels[(size_t)b * Umax + j - 1], V));
    uint4* dst = reinterpret_cast<uint4*>(wy + ((size_t)b * Up + j) * He);
    const uint4* src = reinterpret_cast<const uint4*>(w_bf16 + (size_t)row * He);
    for (int i = threadIdx.x; i < He / 8; i += blockDim.x) dst[i] = live ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) bias_y[(size_t)b * Up + j] = live ? b_out[row] : 0.f;
}

namespace wm = nvcuda::wmma;
using FragA = wm::fragment<wm::matrix_a, 16, 16, 16, __nv_bfloat16, wm::row_major>;
using FragAT = wm::fragment<wm::matrix_a, 16, 16, 16, __nv_bfloat16, wm::col_major>;
using FragBT = wm::fragment<wm::matrix_b, 16, 16, 16, __nv_bfloat16, wm::col_major>;
using FragB = wm::fragment<wm::matrix_b, 16, 16, 16, __nv_bfloat16, wm::row_major>;
using FragC = wm::fragment<wm::accumulator, 16, 16, 16, float>;

constexpr int kHeadWarps = 4;
constexpr int kHeadNT = 8;   // 16-wide output tiles a warp keeps in registers

// Emissions: C[frames x Up] = e_bf[frames x He] wy[b]^T + bias_y - lse, staged as the lattices' input.
// Block = (16-frame tile, utterance); the 4 warps take the 16-wide label tiles round-robin (short dependent chains:
// the kernel is bound by the latency of its fragment loads, not by their volume); warp-level bf16 MMAs on the
// tensor-core path's own operand roundings, fp32 accumulation.
__global__ void __launch_bounds__(kHeadWarps * 32)
head_emission_kernel(const __nv_bfloat16* __restrict__ e_bf, const __nv_bfloat16* __restrict__ wy,
                     const float* __restrict__ bias_y, const float* __restrict__ lse, const int* __restrict__ tlen32,
                     const long long* __restrict__ ulen, int T, int He, int Umax, int Up, float* __restrict__ emis,
                     float* __restrict__ lp_a, float* __restrict__ lp_b) {
    __shared__ __align__(32) float s_c[kHeadWarps][16 * 16];
    __shared__ float s_blank[16];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 16;
    const int T_b = tlen32[b];
    if (t0 >= T_b) return;
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S = 2 * Umax + 1;
    const __nv_bfloat16* arow = e_bf + ((size_t)b * T + t0) * He;
    const __nv_bfloat16* wyb = wy + (size_t)b * Up * He;
    const int ntiles = (U_b + 1 + 15) / 16;
    for (int nt = warp; nt < ntiles; nt += kHeadWarps) {
        FragC acc;
        wm::fill_fragment(acc, 0.f);
        const __nv_bfloat16* brow = wyb + (size_t)nt * 16 * He;
#pragma unroll 4
        for (int k0 = 0; k0 < He; k0 += 16) {
            FragA fa;
            FragBT fb;
            wm::load_matrix_sync(fa, arow + k0, He);
            wm::load_matrix_sync(fb, brow + k0, He);
            wm::mma_sync(acc, fa, fb, acc);
        }
        wm::store_matrix_sync(s_c[warp], acc, 16, wm::mem_row_major);
        __syncwarp();
        for (int e = lane; e < 256; e += 32) {
            const int f = e >> 4, j = nt * 16 + (e & 15);
            if (t0 + f >= T_b || j > U_b) continue;
            const size_t r = (size_t)b * T + t0 + f;
            const float v = s_c[warp][e] + bias_y[(size_t)b * Up + j] - lse[r];
            emis[r * (Umax + 1) + j] = v;
            if (j > 0) {
                lp_a[r * S + 2 * j - 1] = v;
                if (lp_b) lp_b[r * S + 2 * j - 1] = v;
            } else {
                s_blank[f] = v;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // blank states (even s): the emission of column 0
    for (int e = threadIdx.x; e < 16 * (U_b + 1); e += blockDim.x) {
        const int f = e / (U_b + 1), u = e - f * (U_b + 1);
        if (t0 + f >= T_b) continue;
        const size_t r = (size_t)b * T + t0 + f;
        const float v = s_blank[f];
        lp_a[r * S + 2 * u] = v;
        if (lp_b) lp_b[r * S + 2 * u] = v;
    }
}

// g * occupancy per (frame, {blank, label u}) as bf16 [B][Tp][Up], Tp = T rounded up to 16 (zeros for padded frames /
// columns, so that the 16-frame MMA steps of the consumers never mix utterances), and the bias gradient of the sparse
// part.  Block = (tile of kHeadFrames frames, utterance).
__global__ void __launch_bounds__(kHeadThreads)
head_occ_kernel(const float* __restrict__ emis, const float* __restrict__ alpha_ws, const float* __restrict__ beta_ws,
                const long long* __restrict__ labels, const int* __restrict__ tlen32, const long long* __restrict__ ulen,
                const float* __restrict__ geff, const float* __restrict__ ll, int T, int Tp, int V, int Umax, int Up,
                int blank, __nv_bfloat16* __restrict__ occg, float* __restrict__ d_b) {
    extern __shared__ float s_occ[];   // [Up][kHeadFrames]
    const int b = blockIdx.y, t0 = blockIdx.x * kHeadFrames;
    const int T_b = tlen32[b];
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S = 2 * Umax + 1, S_b = 2 * U_b + 1;
    const int nf = max(0, min(kHeadFrames, T_b - t0));
    const float g = geff[b], l = ll[b];
    for (int i = threadIdx.x; i < Up * kHeadFrames; i += blockDim.x) s_occ[i] = 0.f;
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
    for (int i = threadIdx.x; i < Up * kHeadFrames; i += blockDim.x) {
        const int f = i / Up, j = i - f * Up;
        if (t0 + f < Tp) occg[((size_t)b * Tp + t0 + f) * Up + j] = __float2bfloat16_rn(s_occ[j * kHeadFrames + f]);
    }
    if (g != 0.f) {
        const long long* y = labels + (size_t)b * Umax;
        for (int j = threadIdx.x; j <= U_b; j += blockDim.x) {
            float sum = 0.f;
#pragma unroll
            for (int f = 0; f < kHeadFrames; ++f) sum += s_occ[j * kHeadFrames + f];
            if (sum != 0.f) atomicAdd(d_b + (j == 0 ? blank : clamp_label(y[j - 1], V)), -sum);
        }
    }
}

// d_eouts[frames x He] = dh (dense part, bf16, from the ring kernel) - occg[frames x Up] wy[b][Up x He].
// Block = (16-frame tile, 256-column slab, utterance); warp = 16 frames x 64 columns.
constexpr int kDeNT = 4;
__global__ void __launch_bounds__(kHeadWarps * 32)
head_deouts_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ occg,
                   const __nv_bfloat16* __restrict__ wy, const int* __restrict__ tlen32,
                   const long long* __restrict__ ulen, const float* __restrict__ geff, int T, int Tp, int He, int Umax,
                   int Up, int tpu, int per_super, float* __restrict__ d_eouts) {
    __shared__ __align__(32) float s_c[kHeadWarps][16 * kDeNT * 16];
    const int b = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 16;
    const int c0 = (blockIdx.y * kHeadWarps + warp) * (kDeNT * 16);
    if (c0 >= He) return;
    const int T_b = tlen32[b];
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    FragC acc[kDeNT];
#pragma unroll
    for (int i = 0; i < kDeNT; ++i) wm::fill_fragment(acc[i], 0.f);
    if (t0 < T_b && geff[b] != 0.f) {
        const __nv_bfloat16* arow = occg + ((size_t)b * Tp + t0) * Up;
        const __nv_bfloat16* wyb = wy + (size_t)b * Up * He + c0;
        for (int k0 = 0; k0 <= U_b; k0 += 16) {
            FragA fa;
            wm::load_matrix_sync(fa, arow + k0, Up);
#pragma unroll
            for (int i = 0; i < kDeNT; ++i) {
                FragB fb;
                wm::load_matrix_sync(fb, wyb + (size_t)k0 * He + i * 16, He);
                wm::mma_sync(acc[i], fa, fb, acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kDeNT; ++i) wm::store_matrix_sync(s_c[warp] + i * 16, acc[i], kDeNT * 16, wm::mem_row_major);
    __syncwarp();
    // lane -> (frame f = lane / 2 .. , 8 consecutive columns): 16-byte dh loads, two 16-byte stores
#pragma unroll
    for (int it = 0; it < (16 * kDeNT * 16) / (32 * 8); ++it) {
        const int e8 = it * 32 + lane;               // index of an 8-column group of the 16 x 64 tile
        const int f = e8 / (kDeNT * 2), cc = (e8 % (kDeNT * 2)) * 8;
        if (t0 + f >= T) continue;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        if (t0 + f < T_b) {
            // tile-major dh with one cell per frame: row = super-utterance * tpu * 128 + frame inside it
            const size_t row = (size_t)(b / per_super) * tpu * kTileM + (size_t)(b % per_super) * T + t0 + f;
            const uint4 dv = *reinterpret_cast<const uint4*>(dh + row * He + c0 + cc);
            const float* sc = s_c[warp] + f * (kDeNT * 16) + cc;
            o0.x = __uint_as_float(dv.x << 16) - sc[0]; o0.y = __uint_as_float(dv.x & 0xffff0000u) - sc[1];
            o0.z = __uint_as_float(dv.y << 16) - sc[2]; o0.w = __uint_as_float(dv.y & 0xffff0000u) - sc[3];
            o1.x = __uint_as_float(dv.z << 16) - sc[4]; o1.y = __uint_as_float(dv.z & 0xffff0000u) - sc[5];
            o1.z = __uint_as_float(dv.w << 16) - sc[6]; o1.w = __uint_as_float(dv.w & 0xffff0000u) - sc[7];
        }
        float4* dst = reinterpret_cast<float4*>(d_eouts + ((size_t)b * T + t0 + f) * He + c0 + cc);
        dst[0] = o0;
        dst[1] = o1;
    }
}

// d_W[row_j,:] -= sum_t occg[t][j] e_bf[b,t,:]:  C[16 j x 128 cols] = occg^T[16 x T_b] e_bf[T_b x 128].
// Block = (16 columns j of the utterance's {blank, labels}, 128-column slab, utterance); the 4 warps split the frames
// and are summed through shared memory before the scatter.
__global__ void __launch_bounds__(kHeadWarps * 32)
head_dw_kernel(const __nv_bfloat16* __restrict__ e_bf, const __nv_bfloat16* __restrict__ occg,
               const long long* __restrict__ labels, const int* __restrict__ tlen32, const long long* __restrict__ ulen,
               const float* __restrict__ geff, int T, int Tp, int He, int V, int Umax, int Up, int blank,
               float* __restrict__ d_w) {
    __shared__ __align__(32) float s_c[kHeadWarps][16 * kHeadNT * 16];
    const int b = blockIdx.z, j0 = blockIdx.x * 16, c0 = blockIdx.y * (kHeadNT * 16);
    const long long U_bl = ulen[b];
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    if (j0 > U_b || geff[b] == 0.f) return;
    const int warp = threadIdx.x >> 5;
    const int T_b = tlen32[b];
    FragC acc[kHeadNT];
#pragma unroll
    for (int i = 0; i < kHeadNT; ++i) wm::fill_fragment(acc[i], 0.f);
    // frames past T_b inside the last 16-frame step: the occg rows are zero there (own padded rows, [B][Tp][Up]); the
    // e_bf rows are zero (padded frames), the next utterance's first frames or the buffer's zeroed tail -- all finite
    for (int k0 = warp * 16; k0 < T_b; k0 += kHeadWarps * 16) {
        FragAT fa;
        wm::load_matrix_sync(fa, occg + ((size_t)b * Tp + k0) * Up + j0, Up);
        const __nv_bfloat16* brow = e_bf + ((size_t)b * T + k0) * He + c0;
#pragma unroll
        for (int i = 0; i < kHeadNT; ++i) {
            FragB fb;
            wm::load_matrix_sync(fb, brow + i * 16, He);
            wm::mma_sync(acc[i], fa, fb, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < kHeadNT; ++i) wm::store_matrix_sync(s_c[warp] + i * 16, acc[i], kHeadNT * 16, wm::mem_row_major);
    __syncthreads();
    const long long* y = labels + (size_t)b * Umax;
    for (int e = threadIdx.x; e < 16 * kHeadNT * 16; e += blockDim.x) {
        const int jj = j0 + e / (kHeadNT * 16), c = c0 + e % (kHeadNT * 16);
        if (jj > U_b) continue;
        const float v = s_c[0][e] + s_c[1][e] + s_c[2][e] + s_c[3][e];
        const int row = jj == 0 ? blank : clamp_label(y[jj - 1], V);
        if (v != 0.f) atomicAdd(d_w + (size_t)row * He + c, -v);
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
    EMO_REQUIRE(joint_ring_supported(head_super(B, T).Bs, head_super(B, T).Ts, 1, He, V) && He % (kHeadNT * 16) == 0,
                EMO_UNSUPPORTED_SHAPE, "ctc_head: unsupported shape (B <= 1024, T < 65536)");
    EMO_REQUIRE((long long)B * T * He < (1ll << 31), EMO_UNSUPPORTED_SHAPE, "ctc_head: eouts exceeds 2^31 elements");
    EMO_REQUIRE(((uintptr_t)eouts & 15) == 0 && ((uintptr_t)w & 15) == 0, EMO_BAD_ARG,
                "ctc_head: eouts / weight must be 16-byte aligned");
    return EMO_OK;
}

int head_casts(const float* eouts, const float* w, const float* b, const long long* labels, const long long* ulen,
               int B, int T, int He, int V, int Umax, int blank, const HeadWs& ws, cudaStream_t st) {
    const int Vp = padded_vocab(V);
    const int Up = (Umax + 1 + 15) / 16 * 16;
    const size_t nw = (size_t)V * He, ne = (size_t)B * T * He;
    f32_to_bf16_kernel<<<ceil_div(nw, 4 * 256), 256, 0, st>>>(w, ws.w_bf16, nw);
    EMO_CHECK_LAUNCH("f32_to_bf16_kernel");
    head_cast_e_kernel<<<ceil_div(ne, 4 * 256), 256, 0, st>>>(eouts, ws.tlen32, T, He, ne, ws.e16, ws.e_bf);
    EMO_CHECK_LAUNCH("head_cast_e_kernel");
    EMO_CUDA(cudaMemsetAsync(ws.e_bf + ne, 0, (size_t)16 * He * sizeof(__nv_bfloat16), st));
    if (Vp != V) {
        const size_t n_tail = (size_t)(Vp - V) * He;
        pad_vocab_kernel<<<ceil_div(max(n_tail, (size_t)Vp), 256), 256, 0, st>>>(ws.w_bf16 + nw, n_tail, b, ws.b_pad, V, Vp);
        EMO_CHECK_LAUNCH("pad_vocab_kernel");
    }
    head_gather_kernel<<<dim3(Up, B), 32, 0, st>>>(ws.w_bf16, b, labels, ulen, He, V, Umax, Up, blank, ws.wy, ws.bias_y);
    EMO_CHECK_LAUNCH("head_gather_kernel");
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
    rc = head_casts(eouts, w, b, labels, ulen, B, T, He, V, Umax, blank, L, st);
    if (rc) return rc;
    const int Vp = padded_vocab(V);
    const float* bias = Vp != V ? L.b_pad : b;
    // dense part: lse[b,t] and the blank log-prob (lp2[..., 0]); labels are unused with one cell per frame
    // (padded frames are ordinary rows here: their lse / blank log-prob are computed and never read)
    rc = joint_fwd_launch(L.w_bf16, L.e16, nullptr, bias, L.ulen32, L.tsup32, L.ulen32, su.Bs, su.Ts, 1, He, Vp, blank,
                          L.lp2, lse, 1, st);
    if (rc) return rc;
    const int Up = (Umax + 1 + 15) / 16 * 16;
    head_emission_kernel<<<dim3(ceil_div(T, 16), B), kHeadWarps * 32, 0, st>>>(
        L.e_bf, L.wy, L.bias_y, lse, L.tlen32, ulen, T, He, Umax, Up, emis, alpha_ws, beta_ws);
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
    rc = head_casts(eouts, w, b, labels, ulen, B, T, He, V, Umax, blank, L, st);
    if (rc) return rc;
    const int Vp = padded_vocab(V);
    const float* bias = Vp != V ? L.b_pad : b;
    EMO_CUDA(cudaMemsetAsync(d_w, 0, (size_t)V * He * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b, 0, (size_t)V * sizeof(float), st));
    // dense part: dz = g softmax(z) -> d_W, d_b, dh (bf16, tile-major)
    rc = joint_bwd_ring_launch(L.w_bf16, L.e16, nullptr, bias, L.ulen32, L.tsup32, L.ulen32, lse, nullptr, nullptr,
                               L.grow, nullptr, su.Bs, su.Ts, 1, He, Vp, V, blank, 1, L.dh, L.ring, d_w, d_b, st);
    if (rc) return rc;
    // sparse part: the entries of the blank-extended label sequence
    const int Up = (Umax + 1 + 15) / 16 * 16, Tp = (T + 15) / 16 * 16;
    head_occ_kernel<<<dim3(Tp / kHeadFrames, B), kHeadThreads, (size_t)Up * kHeadFrames * sizeof(float), st>>>(
        emis, alpha_ws, beta_ws, labels, L.tlen32, ulen, L.geff, L.ll, T, Tp, V, Umax, Up, blank, L.occg, d_b);
    EMO_CHECK_LAUNCH("head_occ_kernel");
    head_deouts_kernel<<<dim3(ceil_div(T, 16), ceil_div(He, kHeadWarps * kDeNT * 16), B), kHeadWarps * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(L.dh), L.occg, L.wy, L.tlen32, ulen, L.geff, T, Tp, He, Umax, Up,
        tiles128_per_utt(su.Ts, 1), B / su.Bs, d_eouts);
    EMO_CHECK_LAUNCH("head_deouts_kernel");
    head_dw_kernel<<<dim3(Up / 16, He / (kHeadNT * 16), B), kHeadWarps * 32, 0, st>>>(
        L.e_bf, L.occg, labels, L.tlen32, ulen, L.geff, T, Tp, He, V, Umax, Up, blank, d_w);
    EMO_CHECK_LAUNCH("head_dw_kernel");
    return EMO_OK;
}

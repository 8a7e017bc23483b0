// Projection side of the joint folded into the library:
//   enc_proj = w_enc(eouts) + b,  dec_proj = w_dec(douts) + b          (rnn_transducer.py:57-58,153)
// and their backward (d_eouts, d_douts, d_w_enc, d_b_enc, d_w_dec, d_b_dec), so that a training step of the path is
// two C calls working from the encoder / prediction-network outputs: the projected streams exist only as the fp16
// copies the joint kernels gather from (no fp32 (B,T,J) round trip, no separate cast launches, no autograd glue).
//
// One small tcgen05 GEMM kernel, three operand forms (bf16 operands by TMA, fp32 accumulation in TMEM, one
// 128 x <=256 output tile per CTA, 4-stage mbarrier ring; warp 0 = TMA, warp 1 = MMA issuer, warps 2-5 = epilogue):
//   F  C[m,n] = sum_k A[m,k] B[n,k]      A, B K-major            forward:  x W^T (+ bias) -> fp16
//   G  C[m,n] = sum_k A[m,k] B[k,n]      A K-major, B MN-major   d_x = d_proj W        -> fp32
//   H  C[m,n] = sum_k A[k,m] B[k,n]      A, B MN-major           d_W = d_proj^T x      -> fp32, split over k (red.add)
// The MN-major forms are the ones the ring kernel's dh / dW roles use (joint_bwd_ring.cu).
#include "joint_tc.cuh"

namespace emo {
namespace {

constexpr int kPStages = 2;      // 96 KiB of stages: two CTAs per SM, so the 178 tiles of the cfg-3 forward run as one wave
constexpr int kPNT = 256;                 // output columns per CTA
constexpr int kPThreads = 192;
constexpr int kPABytes = kTileM * kBlockK * 2;     // 16 KiB
constexpr int kPBBytes = kPNT * kBlockK * 2;       // 32 KiB
constexpr int kPBox = 8192;                        // [64 x 64] bf16 box
constexpr uint32_t kPDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024, version 1, SW128

enum { FORM_F = 0, FORM_G = 1, FORM_H = 2 };
enum { EPI_F16_BIAS = 0, EPI_F32 = 1, EPI_F32_ADD = 2 };

struct ProjArgs {
    int M, N, K;          // C is M x N; K = contraction length
    int kchunk;           // k-blocks (of 64) per CTA of one output tile (split over K)
    int mt, nt, kt;       // CTAs along M, N and K
    const float* bias;    // EPI_F16_BIAS: (N)
    void* out;            // row-major (M, N): __half (EPI_F16_BIAS) or float
};
struct ProjArgs2 {        // the enc and the dec instance of a form share one launch
    ProjArgs p[2];
};

struct __align__(16) ProjBars {
    uint64_t full[kPStages], empty[kPStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint64_t pdesc(uint32_t lo) { return ((uint64_t)kPDescHi << 32) | lo; }

template <int FORM, int EPI>
__global__ void __launch_bounds__(kPThreads, 2)
proj_gemm_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_b0,
                 const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_b1,
                 const ProjArgs2 args) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int n0ctas = args.p[0].mt * args.p[0].nt * args.p[0].kt;
    const bool second = (int)blockIdx.x >= n0ctas;
    const ProjArgs& a = args.p[second ? 1 : 0];
    const CUtensorMap& tm_a = second ? tm_a1 : tm_a0;
    const CUtensorMap& tm_b = second ? tm_b1 : tm_b0;
    const int cta = (int)blockIdx.x - (second ? n0ctas : 0);
    const int bx = cta % a.mt, by = (cta / a.mt) % a.nt, bz = cta / (a.mt * a.nt);
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)kPStages * kPABytes;
    ProjBars* bars = reinterpret_cast<ProjBars*>(sB + (size_t)kPStages * kPBBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = bx * kTileM, n0 = by * kPNT;
    const int nt = min(kPNT, a.N - n0);                     // live columns of this tile (multiple of 16)
    const int nkb_all = (a.K + kBlockK - 1) / kBlockK;
    const int kb0 = bz * a.kchunk, kb1 = min(nkb_all, kb0 + a.kchunk);
    if (kb0 >= kb1) return;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kPStages; ++i) {
            mbar_init(smem_u32(&bars->full[i]), 1);
            mbar_init(smem_u32(&bars->empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(&bars->tmem_base), kPNT);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int nboxes_b = (nt + 63) / 64;                    // MN-major B: [64 n x 64 k] boxes of the stage
    const uint32_t bytes_a = kPABytes;
    // (a TMA box is always transferred whole: out-of-bounds rows arrive as zeros)
    const uint32_t bytes_b = FORM == FORM_F ? (uint32_t)min(a.N, kPNT) * kBlockK * 2 : (uint32_t)nboxes_b * kPBox;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bars->full[s]);
                mbar_arrive_expect_tx(full, bytes_a + bytes_b);
                const uint32_t da = smem_u32(sA + (size_t)s * kPABytes), db = smem_u32(sB + (size_t)s * kPBBytes);
                if (FORM == FORM_H) {       // A = [k rows x m cols]: two [64 m x 64 k] boxes
                    tma_load_2d(da, &tm_a, m0, kb * kBlockK, full);
                    tma_load_2d(da + kPBox, &tm_a, m0 + 64, kb * kBlockK, full);
                } else {                    // A = [m rows x k cols]: one [64 k x 128 m] box
                    tma_load_2d(da, &tm_a, kb * kBlockK, m0, full);
                }
                if (FORM == FORM_F) {       // B = [n rows x k cols]: one [64 k x nt n] box (tensor map built per launch)
                    tma_load_2d(db, &tm_b, kb * kBlockK, n0, full);
                } else {                    // B = [k rows x n cols]: [64 n x 64 k] boxes
                    for (int i = 0; i < nboxes_b; ++i) tma_load_2d(db + i * kPBox, &tm_b, n0 + i * 64, kb * kBlockK, full);
                }
                if (++s == kPStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        uint32_t s = 0, ph = 0;
        const uint32_t idesc = umma_idesc_bf16(kTileM, (nt + 15) / 16 * 16, FORM == FORM_H ? 1 : 0, FORM == FORM_F ? 0 : 1);
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&bars->full[s]), ph);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t a_addr = smem_u32(sA + (size_t)s * kPABytes), b_addr = smem_u32(sB + (size_t)s * kPBBytes);
#pragma unroll
                for (int k16 = 0; k16 < kBlockK / 16; ++k16) {
                    // K-major: advance 32 bytes inside the 128-byte swizzle row; MN-major: 16 k-rows = 2048 bytes
                    const uint32_t alo = FORM == FORM_H ? (((a_addr & 0x3FFFFu) >> 4) | ((kPBox >> 4) << 16)) + k16 * (2048 >> 4)
                                                        : (((a_addr & 0x3FFFFu) >> 4) | (1u << 16)) + 2 * k16;
                    const uint32_t blo = FORM == FORM_F ? (((b_addr & 0x3FFFFu) >> 4) | (1u << 16)) + 2 * k16
                                                        : (((b_addr & 0x3FFFFu) >> 4) | ((kPBox >> 4) << 16)) + k16 * (2048 >> 4);
                    umma_bf16(tmem_base, pdesc(alo), pdesc(blo), idesc, (kb > kb0 || k16 > 0) ? 1u : 0u);
                }
                umma_commit(smem_u32(&bars->empty[s]));
                if (kb == kb1 - 1) umma_commit(smem_u32(&bars->acc_full));
            }
            __syncwarp();
            if (++s == kPStages) { s = 0; ph ^= 1; }
        }
    } else {
        // epilogue: warp w reads TMEM lanes of quadrant (w & 3); thread == output row
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        mbar_wait(smem_u32(&bars->acc_full), 0);
        tc_fence_after();
        for (int c = 0; c < nt; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + c, r);
            tmem_wait_ld();
            if (row >= a.M) continue;
            const int ncol = min(32, nt - c);
            if (EPI == EPI_F16_BIAS) {
                __half* dst = reinterpret_cast<__half*>(a.out) + (size_t)row * a.N + n0 + c;
                for (int i = 0; i < ncol; i += 8) {
                    const float4 b0 = *reinterpret_cast<const float4*>(a.bias + n0 + c + i);
                    const float4 b1 = *reinterpret_cast<const float4*>(a.bias + n0 + c + i + 4);
                    uint4 o;
                    o.x = pack_f16x2(__uint_as_float(r[i]) + b0.x, __uint_as_float(r[i + 1]) + b0.y);
                    o.y = pack_f16x2(__uint_as_float(r[i + 2]) + b0.z, __uint_as_float(r[i + 3]) + b0.w);
                    o.z = pack_f16x2(__uint_as_float(r[i + 4]) + b1.x, __uint_as_float(r[i + 5]) + b1.y);
                    o.w = pack_f16x2(__uint_as_float(r[i + 6]) + b1.z, __uint_as_float(r[i + 7]) + b1.w);
                    *reinterpret_cast<uint4*>(dst + i) = o;
                }
            } else {
                float* dst = reinterpret_cast<float*>(a.out) + (size_t)row * a.N + n0 + c;
                for (int i = 0; i < ncol; i += 4) {
                    if (EPI == EPI_F32) {
                        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                         __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                    } else {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(r[i])),
                                     "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])),
                                     "f"(__uint_as_float(r[i + 3]))
                                     : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kPNT);
    }
}

struct ProjProblem {
    const void* A;
    const void* B;
    int M, N, K;
    const float* bias;
    void* out;
    int ksplit;
};

template <int FORM, int EPI>
int proj_gemm2(const ProjProblem& q0, const ProjProblem& q1, cudaStream_t st) {
    CUtensorMap tm[4];
    ProjArgs2 args;
    int total = 0;
    for (int i = 0; i < 2; ++i) {
        const ProjProblem& q = i ? q1 : q0;
        int rc;
        // tensor maps: (inner, outer, box_inner, box_outer)
        if (FORM == FORM_H) rc = make_tmap_bf16_2d(&tm[2 * i], q.A, (uint64_t)q.M, (uint64_t)q.K, 64, 64);
        else rc = make_tmap_bf16_2d(&tm[2 * i], q.A, (uint64_t)q.K, (uint64_t)q.M, kBlockK, kTileM);
        if (rc) return rc;
        if (FORM == FORM_F) rc = make_tmap_bf16_2d(&tm[2 * i + 1], q.B, (uint64_t)q.K, (uint64_t)q.N, kBlockK, min(q.N, kPNT));
        else rc = make_tmap_bf16_2d(&tm[2 * i + 1], q.B, (uint64_t)q.N, (uint64_t)q.K, 64, 64);
        if (rc) return rc;
        ProjArgs& a = args.p[i];
        a.M = q.M; a.N = q.N; a.K = q.K; a.bias = q.bias; a.out = q.out;
        const int nkb = ceil_div(q.K, kBlockK);
        const int ks = max(1, min(q.ksplit, nkb));
        a.kchunk = ceil_div(nkb, ks);
        a.mt = ceil_div(q.M, kTileM); a.nt = ceil_div(q.N, kPNT); a.kt = ceil_div(nkb, a.kchunk);
        total += a.mt * a.nt * a.kt;
    }
    const size_t smem = (size_t)kPStages * (kPABytes + kPBBytes) + sizeof(ProjBars);
    auto kern = proj_gemm_kernel<FORM, EPI>;
    EMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<total, kPThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], args);
    EMO_CHECK_LAUNCH("proj_gemm_kernel");
    return EMO_OK;
}

// every fp32 -> bf16 cast of a call in ONE launch (up to 5 segments)
struct CastSeg {
    const float* src;
    __nv_bfloat16* dst;
    unsigned long long n;
};
struct CastArgs {
    CastSeg seg[5];
    int nseg;
};
__global__ void multi_cast_kernel(const CastArgs a) {
    const CastSeg s = a.seg[blockIdx.y];
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < s.n; i += (size_t)gridDim.x * blockDim.x * 4) {
        if (i + 3 < s.n) {
            const float4 v = *reinterpret_cast<const float4*>(s.src + i);
            *reinterpret_cast<uint2*>(s.dst + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        } else {
            for (size_t j = i; j < s.n; ++j) s.dst[j] = __float2bfloat16_rn(s.src[j]);
        }
    }
}

// bf16 copy + column sums (bias gradient) of an (M, N) fp32 matrix in one pass: block = 32 columns x 8 row lanes
__global__ void cast_colsum_kernel(const float* __restrict__ x, int M, int N, __nv_bfloat16* __restrict__ xb,
                                   float* __restrict__ out, float* __restrict__ out2) {
    __shared__ float s[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < N)
        for (int r = blockIdx.y * 8 + threadIdx.y; r < M; r += gridDim.y * 8) {
            const float v = x[(size_t)r * N + c];
            xb[(size_t)r * N + c] = __float2bfloat16_rn(v);
            acc += v;
        }
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
        atomicAdd(out + c, t);
        atomicAdd(out2 + c, t);
    }
}

struct FullWs {   // forward workspace, kept by the caller for the backward
    __nv_bfloat16* w_out_bf;   // (Vp, J)
    __half* enc16;             // (B, T, J)
    __half* dec16;             // (B, U1, J)
    float* b_pad;              // (Vp)
    __nv_bfloat16* e_bf;       // (B*T, He)
    __nv_bfloat16* d_bf;       // (B*U1, Hd)
    __nv_bfloat16* wenc_bf;    // (J, He)
    __nv_bfloat16* wdec_bf;    // (J, Hd)
    size_t total;
};
FullWs full_fwd_layout(void* base, int B, int T, int U1, int He, int Hd, int J, int V) {
    FullWs w;
    const size_t Vp = (size_t)padded_vocab(V);
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 1024); return r; };
    w.w_out_bf = reinterpret_cast<__nv_bfloat16*>(take(Vp * J * 2));
    w.enc16 = reinterpret_cast<__half*>(take((size_t)B * T * J * 2));
    w.dec16 = reinterpret_cast<__half*>(take((size_t)B * U1 * J * 2));
    w.b_pad = reinterpret_cast<float*>(take(Vp * 4));
    w.e_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * T * He * 2));
    w.d_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * U1 * Hd * 2));
    w.wenc_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)J * He * 2));
    w.wdec_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)J * Hd * 2));
    w.total = off;
    return w;
}
struct FullBwdWs {
    void* dh;                  // tile-major bf16 dh
    void* ring;
    float* d_enc;              // (B*T, J) fp32
    float* d_dec;              // (B*U1, J) fp32 (accumulated with red.add)
    __nv_bfloat16* d_enc_bf;
    __nv_bfloat16* d_dec_bf;
    size_t total;
};
FullBwdWs full_bwd_layout(void* base, int B, int T, int U1, int J, int V) {
    FullBwdWs w;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* r = p + off; off += align_up(bytes, 1024); return r; };
    w.dh = take(dh_bytes_for(B, T, U1, J));
    w.ring = take(joint_ring_workspace(B, T, U1, J, V));
    w.d_enc = reinterpret_cast<float*>(take((size_t)B * T * J * 4));
    w.d_dec = reinterpret_cast<float*>(take((size_t)B * U1 * J * 4));
    w.d_enc_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * T * J * 2));
    w.d_dec_bf = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * U1 * J * 2));
    w.total = off;
    return w;
}

int full_check(int B, int T, int U1, int He, int Hd, int J, int V, int blank) {
    int rc = check_bf16_shape(B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(He > 0 && Hd > 0 && He % 16 == 0 && Hd % 16 == 0, EMO_UNSUPPORTED_SHAPE,
                "joint_full: enc / dec hidden sizes (%d, %d) must be multiples of 16", He, Hd);
    EMO_REQUIRE(joint_ring_supported(B, T, U1, J, V), EMO_UNSUPPORTED_SHAPE, "joint_full: unsupported joint shape");
    return EMO_OK;
}

}  // namespace
}  // namespace emo

using namespace emo;

extern "C" int emo_rnnt_joint_full_supported(int B, int T, int U1, int He, int Hd, int J, int V) {
    if (B <= 0 || T <= 0 || U1 <= 0 || He <= 0 || Hd <= 0 || J <= 0 || V <= 0) return 0;
    if (He % 16 || Hd % 16 || J % 128 || J > kMaxKBlocks * kBlockK) return 0;
    if ((long long)B * T * J >= (1ll << 31) || (long long)B * U1 * J >= (1ll << 31)) return 0;
    return joint_ring_supported(B, T, U1, J, V) ? 1 : 0;
}

extern "C" size_t emo_rnnt_joint_full_workspace_bytes(int op, int B, int T, int U1, int He, int Hd, int J, int V) {
    if (!emo_rnnt_joint_full_supported(B, T, U1, He, Hd, J, V)) return 0;
    if (op == 0) return full_fwd_layout(nullptr, B, T, U1, He, Hd, J, V).total;
    if (op == 1) return full_bwd_layout(nullptr, B, T, U1, J, V).total;
    return 0;
}

extern "C" int emo_rnnt_joint_full_fwd(const float* eouts, const float* douts, const float* w_enc, const float* b_enc,
                                       const float* w_dec, const float* b_dec, const float* w_out, const float* b_out,
                                       const int* labels, const int* tlen, const int* ulen, int B, int T, int U1, int He,
                                       int Hd, int J, int V, int blank, float* lp2, float* lse, void* fws,
                                       size_t fws_bytes, void* stream) {
    EMO_REQUIRE(eouts && douts && w_enc && b_enc && w_dec && b_dec && w_out && b_out && labels && tlen && ulen && lp2 &&
                    lse && fws, EMO_BAD_ARG, "joint_full_fwd: null pointer");
    int rc = full_check(B, T, U1, He, Hd, J, V, blank);
    if (rc) return rc;
    const FullWs L = full_fwd_layout(fws, B, T, U1, He, Hd, J, V);
    EMO_REQUIRE(fws_bytes >= L.total && ((uintptr_t)fws & 255) == 0, EMO_WORKSPACE_TOO_SMALL,
                "joint_full_fwd: workspace too small or not 256-byte aligned");
    EMO_REQUIRE((((uintptr_t)eouts | (uintptr_t)douts | (uintptr_t)w_enc | (uintptr_t)w_dec | (uintptr_t)w_out |
                  (uintptr_t)b_enc | (uintptr_t)b_dec) & 15) == 0, EMO_BAD_ARG, "joint_full_fwd: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int Vp = padded_vocab(V);
    CastArgs c;
    c.nseg = 5;
    c.seg[0] = {w_out, L.w_out_bf, (unsigned long long)V * J};
    c.seg[1] = {eouts, L.e_bf, (unsigned long long)B * T * He};
    c.seg[2] = {douts, L.d_bf, (unsigned long long)B * U1 * Hd};
    c.seg[3] = {w_enc, L.wenc_bf, (unsigned long long)J * He};
    c.seg[4] = {w_dec, L.wdec_bf, (unsigned long long)J * Hd};
    multi_cast_kernel<<<dim3(2 * sm_count(), 5), 256, 0, st>>>(c);
    EMO_CHECK_LAUNCH("multi_cast_kernel");
    const float* bias = b_out;
    if (Vp != V) {
        const size_t n_tail = (size_t)(Vp - V) * J;
        pad_vocab_kernel<<<ceil_div(max(n_tail, (size_t)Vp), 256), 256, 0, st>>>(L.w_out_bf + (size_t)V * J, n_tail, b_out,
                                                                                L.b_pad, V, Vp);
        EMO_CHECK_LAUNCH("pad_vocab_kernel");
        bias = L.b_pad;
    }
    // projections straight to the fp16 streams the joint kernels gather from
    const ProjProblem pe = {L.e_bf, L.wenc_bf, B * T, J, He, b_enc, L.enc16, 1};
    const ProjProblem pd = {L.d_bf, L.wdec_bf, B * U1, J, Hd, b_dec, L.dec16, 1};
    if ((rc = proj_gemm2<FORM_F, EPI_F16_BIAS>(pe, pd, st))) return rc;
    return joint_fwd_launch(L.w_out_bf, L.enc16, L.dec16, bias, labels, tlen, ulen, B, T, U1, J, Vp, blank, lp2, lse, 0, st);
}

extern "C" int emo_rnnt_joint_full_bwd(const float* b_out, const int* labels, const int* tlen, const int* ulen,
                                       const float* lse, const float* lp2, const float* gamma2, const float* grad_cost,
                                       const float* grad_lse, const void* fws, int B, int T, int U1, int He, int Hd, int J,
                                       int V, int blank, float* d_eouts, float* d_douts, float* d_w_enc, float* d_b_enc,
                                       float* d_w_dec, float* d_b_dec, float* d_w_out, float* d_b_out, void* ws,
                                       size_t ws_bytes, void* stream) {
    EMO_REQUIRE(b_out && labels && tlen && ulen && lse && lp2 && gamma2 && grad_cost && fws && d_eouts && d_douts &&
                    d_w_enc && d_b_enc && d_w_dec && d_b_dec && d_w_out && d_b_out && ws, EMO_BAD_ARG,
                "joint_full_bwd: null pointer");
    int rc = full_check(B, T, U1, He, Hd, J, V, blank);
    if (rc) return rc;
    const FullWs F = full_fwd_layout(const_cast<void*>(fws), B, T, U1, He, Hd, J, V);
    const FullBwdWs L = full_bwd_layout(ws, B, T, U1, J, V);
    EMO_REQUIRE(ws_bytes >= L.total && ((uintptr_t)ws & 255) == 0, EMO_WORKSPACE_TOO_SMALL,
                "joint_full_bwd: workspace too small or not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int Vp = padded_vocab(V);
    const float* bias = Vp != V ? F.b_pad : b_out;
    EMO_CUDA(cudaMemsetAsync(d_w_out, 0, (size_t)V * J * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_out, 0, (size_t)V * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(L.d_dec, 0, (size_t)B * U1 * J * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_w_enc, 0, (size_t)J * He * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_w_dec, 0, (size_t)J * Hd * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_enc, 0, (size_t)J * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_dec, 0, (size_t)J * sizeof(float), st));
    rc = joint_bwd_ring_launch(F.w_out_bf, F.enc16, F.dec16, bias, labels, tlen, ulen, lse, lp2, gamma2, grad_cost,
                               grad_lse, B, T, U1, J, Vp, V, blank, 0, L.dh, L.ring, d_w_out, d_b_out, st);
    if (rc) return rc;
    // axis reductions; the bf16 copy of d_enc_proj (operand of the projection backward) comes out of the same pass
    rc = joint_reduce_dh_launch_f16(L.dh, F.enc16, F.dec16, tlen, ulen, B, T, U1, J, L.d_enc, L.d_dec, L.d_enc_bf, st);
    if (rc) return rc;
    // d_dec_proj was accumulated with atomics over the frame blocks: its bf16 copy needs a second pass, which also forms
    // the bias gradients.  d_b_enc = sum_{b,t} d_enc_proj and d_b_dec = sum_{b,u} d_dec_proj are the SAME vector (both are
    // the sum of dpre over all lattice cells), so one set of column sums of the smaller matrix serves both.
    cast_colsum_kernel<<<dim3(ceil_div(J, 32), 32), dim3(32, 8), 0, st>>>(L.d_dec, B * U1, J, L.d_dec_bf, d_b_dec, d_b_enc);
    EMO_CHECK_LAUNCH("cast_colsum_kernel");
    // d_x = d_proj W   (A K-major, B MN-major);  d_W = d_proj^T x   (both MN-major, split over the rows)
    const ProjProblem ge = {L.d_enc_bf, F.wenc_bf, B * T, He, J, nullptr, d_eouts, 1};
    const ProjProblem gd = {L.d_dec_bf, F.wdec_bf, B * U1, Hd, J, nullptr, d_douts, 1};
    if ((rc = proj_gemm2<FORM_G, EPI_F32>(ge, gd, st))) return rc;
    const ProjProblem he = {L.d_enc_bf, F.e_bf, J, He, B * T, nullptr, d_w_enc, 32};
    const ProjProblem hd = {L.d_dec_bf, F.d_bf, J, Hd, B * U1, nullptr, d_w_dec, 16};
    if ((rc = proj_gemm2<FORM_H, EPI_F32_ADD>(he, hd, st))) return rc;
    return EMO_OK;
}

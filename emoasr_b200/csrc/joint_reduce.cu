// Axis reductions of the fused joint's backward (EMO_PREC_BF16): d_enc_proj = sum_u dh (1 - h^2),
// d_dec_proj = sum_t dh (1 - h^2) from the tile-major bf16 dh the ring kernel's dh role wrote.
#include "joint_tc.cuh"

namespace emo {
namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// Both axis reductions of dpre = dh (1 - h^2) in ONE pass over the tile-major dh (bf16, written by the ring
// kernel's dh role); row of (b,t,u) = b * tiles128 * 128 + t * (U_b+1) + u:
//   d_enc_proj[b,t,j] = sum_{u <= U_b} dpre[b,t,u,j]   (0 for t >= T_b)        written directly
//   d_dec_proj[b,u,j] = sum_{t <  T_b} dpre[b,t,u,j]   (0 for u >  U_b)        pre-zeroed, red.add.v4
// Block = (128-column slice, kRedTG frames, utterance).  Warp w owns the rows u = w mod 8: for each of its u
// it loads the kRedTG frames at once (8-byte loads, next u prefetched), sums them in registers for d_dec
// (one vector red per (u, lane)) and keeps per-frame partials for d_enc, combined across the 8 warps through
// shared memory once at the end.  No barrier in the loop.
__device__ __forceinline__ uint32_t one_minus_sq_f16x2(uint32_t h) {
    uint32_t r;
    asm("{\n\t.reg .b32 nh;\n\tneg.f16x2 nh, %1;\n\tfma.rn.f16x2 %0, nh, %1, %2;\n\t}" : "=r"(r) : "r"(h), "r"(0x3C003C00u));
    return r;
}

constexpr int kRedTG = 4;    // frames per block: 4 keeps the kernel at <= 64 registers (4 blocks per SM; with 8 it ran at 25 % occupancy, latency-bound)
constexpr int kRedWarps = 8;
constexpr int kRedCols = 128;   // columns per block: 4 per lane (one 8-byte load of 4 bf16)
// h is RECOMPUTED instead of read back: tanh.approx.f16x2(f16(enc) + f16(dec)) is the instruction sequence of the
// forward's A producers on the same inputs (the forward then rounds h to bf16 for the MMA; the derivative factor
// 1 - h^2 is taken from the unrounded half-precision value), and
// enc_proj / dec_proj (11 MB at cfg 3) are L2-resident -- the kernel reads 0.83 GB of dh from HBM and nothing
// else of that size.  The enc values of the block's kRedTG frames stay in registers for the whole u loop.
// TIn = float: the caller's fp32 projected streams (rounded to fp16 here, as the forward's cast does); TIn = __half: the
// fp16 streams themselves (emo_rnnt_joint_full_*: the projections never exist in fp32).
template <typename TIn>
__device__ __forceinline__ void load4_f16x2(const TIn* p, uint32_t& lo, uint32_t& hi);
template <>
__device__ __forceinline__ void load4_f16x2<float>(const float* p, uint32_t& lo, uint32_t& hi) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    lo = pack_f16x2(v.x, v.y);
    hi = pack_f16x2(v.z, v.w);
}
template <>
__device__ __forceinline__ void load4_f16x2<__half>(const __half* p, uint32_t& lo, uint32_t& hi) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    lo = v.x;
    hi = v.y;
}

template <typename TIn>
__global__ void __launch_bounds__(kRedWarps * 32, 4)
reduce_dh_tanh_kernel(const __nv_bfloat16* __restrict__ dh, const TIn* __restrict__ enc_proj,
                      const TIn* __restrict__ dec_proj, const int* __restrict__ tlen,
                      const int* __restrict__ ulen, int T, int U1, int J, int tpu, float* __restrict__ d_enc,
                      float* __restrict__ d_dec, __nv_bfloat16* __restrict__ d_enc_bf) {
    __shared__ float4 s_enc[kRedWarps][kRedTG][32];
    const int b = blockIdx.z, j0 = blockIdx.x * kRedCols;
    const int t0 = blockIdx.y * kRedTG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T_b = min(max(tlen[b], 1), T), U1b = min(max(ulen[b], 0), U1 - 1) + 1;
    const int nt = max(0, min(kRedTG, T_b - t0));          // valid frames of this block
    const size_t rs = (size_t)J / 4;                        // row stride in uint2
    const size_t off = (((size_t)b * tpu * kTileM + (size_t)t0 * U1b) * J + j0) / 4 + lane;
    const uint2* dbase = reinterpret_cast<const uint2*>(dh) + off;
    // enc(t0+k, j0 + 4 lane .. +3) as two packed f16x2, fixed for the block
    uint32_t e2[kRedTG][2];
#pragma unroll
    for (int k = 0; k < kRedTG; ++k) {
        e2[k][0] = e2[k][1] = 0u;
        if (k < nt) load4_f16x2<TIn>(enc_proj + ((size_t)b * T + t0 + k) * J + j0 + lane * 4, e2[k][0], e2[k][1]);
    }
    // Every address of the u loop is a pointer that advances by a constant: the first version recomputed the 64-bit
    // row offsets of all five loads and of the red in every iteration (ncu source page: ~200 of the loop's 320
    // instructions were integer address arithmetic; the kernel was issue-bound at 3.7 TB/s).  Loads are unconditional
    // (clamped to rows that exist), frames past T_b are masked with a select.
    float e[kRedTG][4];
#pragma unroll
    for (int k = 0; k < kRedTG; ++k) e[k][0] = e[k][1] = e[k][2] = e[k][3] = 0.f;
    if (nt > 0) {
        const int w0 = min(warp, U1b - 1);
        const uint2* dp[kRedTG];
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) dp[k] = dbase + ((size_t)(k < nt ? k : 0) * U1b + w0) * rs;
        const TIn* decp = dec_proj + ((size_t)b * U1 + w0) * J + j0 + lane * 4;
        float* ddp = d_dec + ((size_t)b * U1 + w0) * J + j0 + lane * 4;
        const size_t ustep = (size_t)kRedWarps * rs, jstep = (size_t)kRedWarps * J;
        uint2 cd[kRedTG], nd[kRedTG];
        uint2 cdec, ndec;
        load4_f16x2<TIn>(decp, cdec.x, cdec.y);
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) cd[k] = __ldg(dp[k]);
        for (int u = warp; u < U1b; u += kRedWarps) {
            const bool more = u + kRedWarps < U1b;      // warp-uniform
            if (more) {
                decp += jstep;
#pragma unroll
                for (int k = 0; k < kRedTG; ++k) dp[k] += ustep;
            }
            load4_f16x2<TIn>(decp, ndec.x, ndec.y);     // the row after the last one is never used: reload the last
#pragma unroll
            for (int k = 0; k < kRedTG; ++k) nd[k] = __ldg(dp[k]);
            const uint32_t d2a = cdec.x, d2b = cdec.y;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < kRedTG; ++k) {
                // 1 - h^2 in packed half precision (h = tanh.approx.f16x2 of the forward's own sum; the factor is in
                // [0,1], 11 mantissa bits), then two FMAs per element
                const float2 ga = unpack_f16x2(one_minus_sq_f16x2(tanh_f16x2(hadd2_u32(e2[k][0], d2a))));
                const float2 gb = unpack_f16x2(one_minus_sq_f16x2(tanh_f16x2(hadd2_u32(e2[k][1], d2b))));
                const uint32_t x = k < nt ? cd[k].x : 0u, y = k < nt ? cd[k].y : 0u;
                const float q0 = __uint_as_float(x << 16), q1 = __uint_as_float(x & 0xffff0000u);
                const float q2 = __uint_as_float(y << 16), q3 = __uint_as_float(y & 0xffff0000u);
                a0 = fmaf(q0, ga.x, a0); a1 = fmaf(q1, ga.y, a1); a2 = fmaf(q2, gb.x, a2); a3 = fmaf(q3, gb.y, a3);
                e[k][0] = fmaf(q0, ga.x, e[k][0]); e[k][1] = fmaf(q1, ga.y, e[k][1]);
                e[k][2] = fmaf(q2, gb.x, e[k][2]); e[k][3] = fmaf(q3, gb.y, e[k][3]);
            }
            red_add_v4(ddp, a0, a1, a2, a3);
            ddp += jstep;
#pragma unroll
            for (int k = 0; k < kRedTG; ++k) cd[k] = nd[k];
            cdec = ndec;
        }
    }
#pragma unroll
    for (int k = 0; k < kRedTG; ++k) s_enc[warp][k][lane] = make_float4(e[k][0], e[k][1], e[k][2], e[k][3]);
    __syncthreads();
    for (int i = threadIdx.x; i < kRedTG * 32; i += blockDim.x) {
        const int k = i >> 5, l = i & 31;
        if (t0 + k >= T) continue;
        float4 a = s_enc[0][k][l];
#pragma unroll
        for (int w = 1; w < kRedWarps; ++w) {
            const float4 x = s_enc[w][k][l];
            a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
        }
        reinterpret_cast<float4*>(d_enc + ((size_t)b * T + t0 + k) * J + j0)[l] = a;
        if (d_enc_bf)   // bf16 copy for the projection backward (emo_rnnt_joint_full_bwd)
            reinterpret_cast<uint2*>(d_enc_bf + ((size_t)b * T + t0 + k) * J + j0)[l] =
                make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    }
}

}  // namespace

// d_enc_proj = sum_u dh (1 - h^2), d_dec_proj = sum_t dh (1 - h^2) from the tile-major bf16 dh (d_dec_proj pre-zeroed)
int joint_reduce_dh_launch(const void* dh_ws, const float* enc_proj, const float* dec_proj, const int* tlen,
                           const int* ulen, int B, int T, int U1, int J, float* d_enc_proj, float* d_dec_proj,
                           cudaStream_t st) {
    EMO_REQUIRE(enc_proj && dec_proj && ((uintptr_t)enc_proj & 15) == 0 && ((uintptr_t)dec_proj & 15) == 0, EMO_BAD_ARG,
                "joint_bwd(bf16): enc_proj / dec_proj must be given and 16-byte aligned");
    reduce_dh_tanh_kernel<float><<<dim3(J / kRedCols, ceil_div(T, kRedTG), B), kRedWarps * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dh_ws), enc_proj, dec_proj, tlen, ulen, T, U1, J,
        tiles128_per_utt(T, U1), d_enc_proj, d_dec_proj, nullptr);
    EMO_CHECK_LAUNCH("reduce_dh_tanh_kernel");
    return EMO_OK;
}

// same from the fp16 streams (emo_rnnt_joint_full_bwd); also leaves a bf16 copy of d_enc_proj
int joint_reduce_dh_launch_f16(const void* dh_ws, const void* enc16, const void* dec16, const int* tlen, const int* ulen,
                               int B, int T, int U1, int J, float* d_enc_proj, float* d_dec_proj, void* d_enc_bf,
                               cudaStream_t st) {
    reduce_dh_tanh_kernel<__half><<<dim3(J / kRedCols, ceil_div(T, kRedTG), B), kRedWarps * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dh_ws), reinterpret_cast<const __half*>(enc16),
        reinterpret_cast<const __half*>(dec16), tlen, ulen, T, U1, J, tiles128_per_utt(T, U1), d_enc_proj, d_dec_proj,
        reinterpret_cast<__nv_bfloat16*>(d_enc_bf));
    EMO_CHECK_LAUNCH("reduce_dh_tanh_kernel");
    return EMO_OK;
}

}  // namespace emo

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
constexpr int kRedTG = 8;
constexpr int kRedWarps = 8;
constexpr int kRedCols = 128;   // columns per block: 4 per lane (one 8-byte load of 4 bf16)
// h is RECOMPUTED instead of read back: h = bf16(tanh.approx.f16x2(f16(enc) + f16(dec))) is the exact instruction
// sequence of the forward's A producers on the same inputs, so it reproduces the forward's h bit for bit, and
// enc_proj / dec_proj (11 MB at cfg 3) are L2-resident -- the kernel reads 0.83 GB of dh from HBM and nothing
// else of that size.  The enc values of the block's kRedTG frames stay in registers for the whole u loop.
__global__ void __launch_bounds__(kRedWarps * 32)
reduce_dh_tanh_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ enc_proj,
                      const float* __restrict__ dec_proj, const int* __restrict__ tlen,
                      const int* __restrict__ ulen, int T, int U1, int J, int tpu, float* __restrict__ d_enc,
                      float* __restrict__ d_dec) {
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
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < nt) v = __ldg(reinterpret_cast<const float4*>(enc_proj + ((size_t)b * T + t0 + k) * J + j0) + lane);
        e2[k][0] = pack_f16x2(v.x, v.y);
        e2[k][1] = pack_f16x2(v.z, v.w);
    }
    const float4* dec4 = reinterpret_cast<const float4*>(dec_proj + (size_t)b * U1 * J + j0) + lane;
    float e[kRedTG][4];
#pragma unroll
    for (int k = 0; k < kRedTG; ++k) e[k][0] = e[k][1] = e[k][2] = e[k][3] = 0.f;
    auto load_u = [&](int u, uint2 (&dv)[kRedTG], float4& dc) {
        const bool uok = u < U1b;
        dc = uok ? __ldg(dec4 + (size_t)u * (J / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < kRedTG; ++k)
            dv[k] = (k < nt && uok) ? __ldg(dbase + ((size_t)k * U1b + u) * rs) : make_uint2(0u, 0u);
    };
    uint2 cd[kRedTG], nd[kRedTG];
    float4 cdec, ndec;
    load_u(warp, cd, cdec);
    for (int u = warp; u < U1b; u += kRedWarps) {
        load_u(u + kRedWarps, nd, ndec);
        const uint32_t d2a = pack_f16x2(cdec.x, cdec.y), d2b = pack_f16x2(cdec.z, cdec.w);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) {
            const float2 ha = unpack_f16x2(tanh_f16x2(hadd2_u32(e2[k][0], d2a)));
            const float2 hb = unpack_f16x2(tanh_f16x2(hadd2_u32(e2[k][1], d2b)));
            const uint32_t pa = pack_bf16x2(ha.x, ha.y), pb = pack_bf16x2(hb.x, hb.y);   // h as the cache holds it
            const float h0 = __uint_as_float(pa << 16), h1 = __uint_as_float(pa & 0xffff0000u);
            const float h2 = __uint_as_float(pb << 16), h3 = __uint_as_float(pb & 0xffff0000u);
            const float f0 = __uint_as_float(cd[k].x << 16) * fmaf(-h0, h0, 1.f);
            const float f1 = __uint_as_float(cd[k].x & 0xffff0000u) * fmaf(-h1, h1, 1.f);
            const float f2 = __uint_as_float(cd[k].y << 16) * fmaf(-h2, h2, 1.f);
            const float f3 = __uint_as_float(cd[k].y & 0xffff0000u) * fmaf(-h3, h3, 1.f);
            a0 += f0; a1 += f1; a2 += f2; a3 += f3;
            e[k][0] += f0; e[k][1] += f1; e[k][2] += f2; e[k][3] += f3;
        }
        if (nt > 0) red_add_v4(d_dec + ((size_t)b * U1 + u) * J + j0 + lane * 4, a0, a1, a2, a3);
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) cd[k] = nd[k];
        cdec = ndec;
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
    }
}

}  // namespace

// d_enc_proj = sum_u dh (1 - h^2), d_dec_proj = sum_t dh (1 - h^2) from the tile-major bf16 dh (d_dec_proj pre-zeroed)
int joint_reduce_dh_launch(const void* dh_ws, const float* enc_proj, const float* dec_proj, const int* tlen,
                           const int* ulen, int B, int T, int U1, int J, float* d_enc_proj, float* d_dec_proj,
                           cudaStream_t st) {
    EMO_REQUIRE(enc_proj && dec_proj && ((uintptr_t)enc_proj & 15) == 0 && ((uintptr_t)dec_proj & 15) == 0, EMO_BAD_ARG,
                "joint_bwd(bf16): enc_proj / dec_proj must be given and 16-byte aligned");
    reduce_dh_tanh_kernel<<<dim3(J / kRedCols, ceil_div(T, kRedTG), B), kRedWarps * 32, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dh_ws), enc_proj, dec_proj, tlen, ulen, T, U1, J,
        tiles128_per_utt(T, U1), d_enc_proj, d_dec_proj);
    EMO_CHECK_LAUNCH("reduce_dh_tanh_kernel");
    return EMO_OK;
}

}  // namespace emo

// Shared pieces of the tcgen05 joint kernels (forward: joint_bf16.cu, backward: joint_bwd_ring.cu, joint_reduce.cu).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace emo {
namespace {

using namespace tc;

constexpr int kTileM = 128;          // cells per tile
constexpr int kBlockK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int kChunkN = 256;         // vocab columns per accumulator buffer
constexpr int kMaxKBlocks = 8;       // J <= 512
constexpr int kABlockBytes = kTileM * kBlockK * 2;    // 16 KiB
constexpr int kThreads = 384;
constexpr int kSmemLimit = 232448;                    // 227 KiB opt-in maximum per CTA
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// the tensor-core kernels work on a vocabulary padded to a multiple of 32 (joint_bf16_casts)
__host__ __device__ inline int padded_vocab(int V) { return (V + 31) / 32 * 32; }

struct TileInfo {
    int b, first_cell, n_cells, U1b;  // n_cells = valid cells of the utterance
};

// dh workspace of the backward (bf16, tile-major): layout [B * tiles128_per_utt * 128, J], row = (b *
// tiles128_per_utt + first_cell / 128) * 128 + row-in-tile for the flattened valid cell m = t * (U_b+1) + u;
// tiles128_per_utt is even so a CTA pair always owns two slots.
__host__ __device__ inline int tiles128_per_utt(int T, int U1) { return 2 * ((T * U1 + 255) / 256); }
inline size_t dh_bytes_for(int B, int T, int U1, int J) {
    return (size_t)B * tiles128_per_utt(T, U1) * kTileM * J * sizeof(__nv_bfloat16);
}

// kCtas = 2: a CTA pair (cluster of 2, cta_group::2) per 256-cell tile, 128 cells per CTA; `rank` selects this
// CTA's half.  Both CTAs of a pair get the same answer.  (kCtas = 1: one CTA per 128-cell tile.)
template <int kCtas>
__device__ __forceinline__ bool tile_info(int tile, int tiles_per_utt, uint32_t rank, const int* tlen,
                                          const int* ulen, int T, int U1, TileInfo& ti) {
    ti.b = tile / tiles_per_utt;
    int i = tile - ti.b * tiles_per_utt;
    int T_b = min(max(__ldg(tlen + ti.b), 1), T);
    ti.U1b = min(max(__ldg(ulen + ti.b), 0), U1 - 1) + 1;
    ti.n_cells = T_b * ti.U1b;
    ti.first_cell = i * (kCtas * kTileM) + (int)rank * kTileM;
    return i * (kCtas * kTileM) < ti.n_cells;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        uint2 o = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (; i < n; ++i) dst[i] = __float2half_rn(src[i]);
    }
}

// Vocabulary sizes that are not a multiple of 32 (e.g. the reference's 10872 / 9798 SentencePiece vocabularies) run
// on a padded copy: pad rows of the bf16 w_out are zero and the pad entries of the bias copy are -1e30, so the pad
// logits contribute exp2(-huge) = 0 to every log-sum-exp and get dz = 0.
__global__ void pad_vocab_kernel(__nv_bfloat16* __restrict__ w_tail, size_t n_tail, const float* __restrict__ b_out,
                                 float* __restrict__ b_pad, int V, int Vp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tail) w_tail[i] = __float2bfloat16_rn(0.f);
    if (i < (size_t)Vp) b_pad[i] = i < (size_t)V ? b_out[i] : -1e30f;
}

// x[d] for a per-thread d in [0,32) without dynamic register indexing: five select levels
__device__ __forceinline__ float mux32(const float (&x)[32], int d) {
    float a[16], b[8], c[4], e[2];
    const bool s4 = d & 16, s3 = d & 8, s2 = d & 4, s1 = d & 2, s0 = d & 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = s4 ? x[i + 16] : x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = s3 ? a[i + 8] : a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = s2 ? b[i + 4] : b[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) e[i] = s1 ? c[i + 2] : c[i];
    return s0 ? e[1] : e[0];
}

// ---- host side ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint32_t box_inner, uint32_t box_outer,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    // libcuda is reached through the runtime (no link-time dependency on libcuda.so.1, so the
    // library also loads on a machine without a driver)
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorName(e));
            return EMO_NO_DEVICE;
        }
        encode = (PFN_encodeTiled)fn;
    }
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
        return EMO_LAUNCH_FAILURE;
    }
    return EMO_OK;
}

int check_bf16_shape(int B, int T, int U1, int J, int V, int blank) {
    EMO_REQUIRE(B > 0 && T > 0 && U1 > 0 && J > 0 && V > 0, EMO_BAD_ARG, "joint(bf16): bad sizes");
    EMO_REQUIRE(blank >= 0 && blank < V, EMO_BAD_ARG, "joint(bf16): blank %d outside [0,%d)", blank, V);
    EMO_REQUIRE(J % 128 == 0 && J <= kMaxKBlocks * kBlockK, EMO_UNSUPPORTED_SHAPE,
                "joint(bf16): joint_hidden_size %d must be a multiple of 128 and <= 512 "
                "(use precision fp32 for other sizes)", J);
    EMO_REQUIRE((long long)B * T * J < (1ll << 31) && (long long)B * U1 * J < (1ll << 31),
                EMO_UNSUPPORTED_SHAPE, "joint(bf16): projected streams exceed 2^31 elements");
    return EMO_OK;
}


}  // namespace
}  // namespace emo

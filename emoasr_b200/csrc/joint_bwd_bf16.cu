// EMO_PREC_BF16 backward of the fused joint (tcgen05 / TMEM / TMA).
//
// The dense gradient of the logits
//     dz[cell,v] = g_b * (gamma * softmax(z)[v] - gamma_blank 1[v=blank] - gamma_label 1[v=label])
// is recomputed tile by tile from (h, w_out, lse) and consumed straight from shared memory by the
// MMAs that need it; it never reaches global memory.  Two kernels share one skeleton:
//
//   kDW = false  "dh kernel", cell-stationary like the forward: for a 128-cell tile and one
//                J-part (<= 256 hidden units) at a time, loop over 128-wide vocab chunks:
//                  z = h W_c^T -> dz (bf16, smem) -> dh_part += dz W_c[:, part]
//                then dpre = dh (1 - h^2) is written as bf16 (B,T,U1,J) for the two axis
//                reductions (d_enc_proj = sum_u, d_dec_proj = sum_t) done by small kernels.
//   kDW = true   "dW kernel", one (vocab chunk, J-part) role per CTA, persistent over cell tiles:
//                  z = h W_c^T -> dz -> dW[c, part] += dz^T h[:, part]      (accumulates in TMEM
//                over ALL tiles of the CTA, flushed once with red.global.add.v4.f32); the same
//                CTAs produce d_b_out = column sums of dz.
//
// Operand layouts: the h tile and the dz tile are stored once, rows = cells, 128-byte rows of 64
// bf16, 128B swizzle.  Read as K-major they feed z = h W^T and dh = dz W; read as MN-major (same
// bytes, different descriptor) they feed dW = dz^T h.  w_out tiles [128 v x 64 j] arrive by TMA and
// are K-major B for z and MN-major B for dh.
// TMEM: columns [0,256) two z buffers of 128, columns [256,512) the dh / dW accumulator.
#include "joint_tc.cuh"

namespace emo {
namespace {

constexpr int kBwdChunk = 128;                 // vocab columns per z chunk
constexpr int kBwdStages = 4;                  // TMA ring
constexpr int kBwdTileBytes = 128 * kBlockK * 2;   // [128 v x 64 j] bf16 = 16 KiB
constexpr int kDzBytes = 2 * kABlockBytes;     // [128 cells x 128 v] bf16 = 32 KiB
constexpr int kPartBlocks = 4;                 // J-part = up to 4 K blocks = 256 hidden units
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024, v1, SW128

struct __align__(16) BwdBarriers {
    uint64_t b_full[kBwdStages], b_empty[kBwdStages];
    uint64_t a_full[kMaxKBlocks], a_empty[kMaxKBlocks];
    uint64_t z_full[2], z_empty[2];
    uint64_t dz_full, dz_empty;
    uint64_t acc2_full, acc2_empty;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
    return ((uint64_t)kDescHiSw128 << 32) | lo;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
// lane l returns sum over the 32 lanes of v[l] (31 shuffles)
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = lane & s;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

struct RowCtx {
    bool valid;
    size_t cell;
    float c2;      // -lse * log2e
    float gg;      // g * (gamma_blank + gamma_label)
    float corr_b;  // g * gamma_blank
    float corr_l;  // g * gamma_label
    int lab;
};

template <bool kDW>
__global__ void __launch_bounds__(kThreads, 1)
joint_bwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ enc,
                 const float* __restrict__ dec, const float* __restrict__ b_out,
                 const int* __restrict__ labels, const int* __restrict__ tlen,
                 const int* __restrict__ ulen, const float* __restrict__ lse,
                 const float* __restrict__ gamma2, const float* __restrict__ grad_cost, int B, int T,
                 int U1, int J, int V, int blank, int num_splits,
                 __nv_bfloat16* __restrict__ dpre_out,   // !kDW: (B,T,U1,J)
                 float* __restrict__ d_w_out,            //  kDW: (V,J), pre-zeroed
                 float* __restrict__ d_b_out) {          //  kDW: (V), pre-zeroed
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = J / kBlockK;
    const int NCH = (V + kBwdChunk - 1) / kBwdChunk;
    const int NPART = (KB + kPartBlocks - 1) / kPartBlocks;
    uint8_t* sA = smem;
    uint8_t* sDz = sA + (size_t)KB * kABlockBytes;
    uint8_t* sB = sDz + kDzBytes;
    BwdBarriers* bars = reinterpret_cast<BwdBarriers*>(sB + (size_t)kBwdStages * kBwdTileBytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);  // [2][kBwdChunk]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_utt = (T * U1 + kTileM - 1) / kTileM;
    const int total_tiles = B * tiles_per_utt;
    // kDW: role = (vocab chunk, J-part), `num_splits` CTAs per role share the cell tiles
    const int role = kDW ? (int)blockIdx.x % (NCH * NPART) : 0;
    const int role_c = role / NPART, role_p = role % NPART;
    const int tile0 = kDW ? (int)blockIdx.x / (NCH * NPART) : (int)blockIdx.x;
    const int tile_stride = kDW ? num_splits : (int)gridDim.x;
    auto part_blocks = [&](int p) { return min(kPartBlocks, KB - p * kPartBlocks); };
    auto chunk_cols = [&](int c) { return min(kBwdChunk, V - c * kBwdChunk); };

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBwdStages; ++i) {
            mbar_init(smem_u32(&bars->b_full[i]), 1);
            mbar_init(smem_u32(&bars->b_empty[i]), 1);
        }
        for (int i = 0; i < kMaxKBlocks; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), 128);
            mbar_init(smem_u32(&bars->a_empty[i]), kDW ? 1 : 129);  // !kDW: + epilogue readers of h
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->z_empty[i]), 128);
        }
        mbar_init(smem_u32(&bars->dz_full), 128);
        mbar_init(smem_u32(&bars->dz_empty), 1);
        mbar_init(smem_u32(&bars->acc2_full), 1);
        mbar_init(smem_u32(&bars->acc2_empty), 128);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w);
    if (warp == 2) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t tmem_acc2 = tmem_base + 2 * kBwdChunk;

    if (warp == 0) {
        // ===================== TMA producer: same order as the MMA issuer consumes =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            auto load = [&](int x, int y) {
                mbar_wait(smem_u32(&bars->b_empty[stage]), phase ^ 1);
                const uint32_t full = smem_u32(&bars->b_full[stage]);
                mbar_arrive_expect_tx(full, kBwdTileBytes);
                tma_load_2d(smem_u32(sB + (size_t)stage * kBwdTileBytes), &tmap_w, x, y, full);
                if (++stage == kBwdStages) { stage = 0; phase ^= 1; }
            };
            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
                if (kDW) {
                    for (int kb = 0; kb < KB; ++kb) load(kb * kBlockK, role_c * kBwdChunk);
                } else {
                    for (int p = 0; p < NPART; ++p) {
                        const int pb = part_blocks(p);
                        for (int c = 0; c <= NCH; ++c) {
                            if (c < NCH)
                                for (int kb = 0; kb < KB; ++kb) load(kb * kBlockK, c * kBwdChunk);
                            if (c > 0)
                                for (int jb = 0; jb < pb; ++jb)
                                    load((p * kPartBlocks + jb) * kBlockK, (c - 1) * kBwdChunk);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0, zc = 0, dc = 0, ac = 0, tl = 0;
        const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16);
        const uint32_t b_lo0 = desc_lo(smem_u32(sB), 16);
        const uint32_t dz_lo0 = desc_lo(smem_u32(sDz), 16);
        // MN-major views (LBO = 16 KiB between 64-element atoms along M/N)
        const uint32_t a_mn_lo0 = desc_lo(smem_u32(sA), kABlockBytes);
        const uint32_t b_mn_lo0 = desc_lo(smem_u32(sB), kBwdTileBytes);
        const uint32_t dz_mn_lo0 = desc_lo(smem_u32(sDz), kABlockBytes);
        auto advance = [&]() { if (++stage == kBwdStages) { stage = 0; phase ^= 1; } };

        // z[128 x n] = h W_c^T into z buffer zc&1
        auto z_mma = [&](int n, bool wait_a, int release_a_mask) {
            const uint32_t zb = zc & 1;
            mbar_wait(smem_u32(&bars->z_empty[zb]), ((zc >> 1) & 1) ^ 1);
            const uint32_t idesc = umma_idesc_bf16(kTileM, n);
            const uint32_t d_tmem = tmem_base + zb * kBwdChunk;
            for (int kb = 0; kb < KB; ++kb) {
                if (wait_a) mbar_wait(smem_u32(&bars->a_full[kb]), tl & 1);
                mbar_wait(smem_u32(&bars->b_full[stage]), phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_lo = a_lo0 + kb * (kABlockBytes >> 4);
                    const uint32_t b_lo = b_lo0 + stage * (kBwdTileBytes >> 4);
#pragma unroll
                    for (int k16 = 0; k16 < kBlockK / 16; ++k16)
                        umma_bf16(d_tmem, mk_desc(a_lo + 2 * k16), mk_desc(b_lo + 2 * k16), idesc,
                                  (kb | k16) != 0);
                    umma_commit(smem_u32(&bars->b_empty[stage]));
                    if ((release_a_mask >> kb) & 1) umma_commit(smem_u32(&bars->a_empty[kb]));
                    if (kb == KB - 1) umma_commit(smem_u32(&bars->z_full[zb]));
                }
                __syncwarp();
                advance();
            }
            ++zc;
        };

        TileInfo ti;
        bool any_tile = false;
        for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            if (kDW) {
                const int pb = part_blocks(role_p);
                const int part_mask = ((1 << pb) - 1) << (role_p * kPartBlocks);
                z_mma(chunk_cols(role_c), true, ((1 << KB) - 1) & ~part_mask);
                // dW[c, part] += dz^T h[:, part]   (M = 128 vocab rows, N = 64 pb, K = 128 cells)
                mbar_wait(smem_u32(&bars->dz_full), dc & 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t idesc = umma_idesc_bf16(kTileM, pb * kBlockK, 1, 1);
                    const uint32_t hb_lo = a_mn_lo0 + (role_p * kPartBlocks) * (kABlockBytes >> 4);
#pragma unroll
                    for (int kk = 0; kk < kTileM / 16; ++kk)
                        umma_bf16(tmem_acc2, mk_desc(dz_mn_lo0 + kk * (2048 >> 4)),
                                  mk_desc(hb_lo + kk * (2048 >> 4)), idesc, (any_tile || kk) ? 1u : 0u);
                    umma_commit(smem_u32(&bars->dz_empty));
                    for (int jb = 0; jb < pb; ++jb)
                        umma_commit(smem_u32(&bars->a_empty[role_p * kPartBlocks + jb]));
                }
                __syncwarp();
                ++dc;
            } else {
                for (int p = 0; p < NPART; ++p) {
                    const int pb = part_blocks(p);
                    for (int c = 0; c <= NCH; ++c) {
                        if (c < NCH) {
                            const bool last_use = (p == NPART - 1) && (c == NCH - 1);
                            z_mma(chunk_cols(c), p == 0 && c == 0, last_use ? (1 << KB) - 1 : 0);
                        }
                        if (c > 0) {
                            // dh_part += dz W_{c-1}[:, part]   (M = 128 cells, N = 64 per tile, K = n)
                            const int n = chunk_cols(c - 1);
                            if (c == 1) mbar_wait(smem_u32(&bars->acc2_empty), (ac & 1) ^ 1);
                            mbar_wait(smem_u32(&bars->dz_full), dc & 1);
                            const uint32_t idesc = umma_idesc_bf16(kTileM, kBlockK, 0, 1);
                            for (int jb = 0; jb < pb; ++jb) {
                                mbar_wait(smem_u32(&bars->b_full[stage]), phase);
                                tc_fence_after();
                                if (elect_one_sync()) {
                                    const uint32_t b_lo = b_mn_lo0 + stage * (kBwdTileBytes >> 4);
                                    for (int kk = 0; kk < n / 16; ++kk)
                                        umma_bf16(tmem_acc2 + jb * kBlockK,
                                                  mk_desc(dz_lo0 + (kk >> 2) * (kABlockBytes >> 4) + (kk & 3) * 2),
                                                  mk_desc(b_lo + kk * (2048 >> 4)), idesc,
                                                  (c > 1 || kk) ? 1u : 0u);
                                    umma_commit(smem_u32(&bars->b_empty[stage]));
                                    if (jb == pb - 1) {
                                        umma_commit(smem_u32(&bars->dz_empty));
                                        if (c == NCH) umma_commit(smem_u32(&bars->acc2_full));
                                    }
                                }
                                __syncwarp();
                                advance();
                            }
                            ++dc;
                        }
                    }
                    ++ac;
                }
            }
            any_tile = true;
            ++tl;
        }
        if (kDW && any_tile) {
            if (elect_one_sync()) umma_commit(smem_u32(&bars->acc2_full));
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int etid = threadIdx.x - 128;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t zc = 0, dc = 0, ac = 0, tl = 0;
        float colsum[kBwdChunk / 32] = {0.f, 0.f, 0.f, 0.f};
        bool any_tile = false;
        TileInfo ti;

        // z chunk -> dz (bf16) into the shared dz tile; returns after signalling dz_full
        auto epi1 = [&](int c, const RowCtx& rc) {
            const uint32_t zb = zc & 1;
            const int n = chunk_cols(c);
            float* bias = s_bias + zb * kBwdChunk;
            if (etid < n) bias[etid] = __ldg(b_out + c * kBwdChunk + etid);
            named_bar_sync(1, 128);
            mbar_wait(smem_u32(&bars->z_full[zb]), (zc >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_base + zb * kBwdChunk;
            bool dz_free = false;
#pragma unroll
            for (int g = 0; g < kBwdChunk / 32; ++g) {
                if (g >= (n >> 5)) break;
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + g * 32, r);
                tmem_wait_ld();
                const int v0 = c * kBwdChunk + g * 32;
                float d[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(bias + g * 32 + i);
                    d[i + 0] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 0]) + bv.x, kLog2e, rc.c2));
                    d[i + 1] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 1]) + bv.y, kLog2e, rc.c2));
                    d[i + 2] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 2]) + bv.z, kLog2e, rc.c2));
                    d[i + 3] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 3]) + bv.w, kLog2e, rc.c2));
                }
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(d[2 * i], d[2 * i + 1]);
                if (!dz_free) {  // the previous dz tile must have been consumed by its MMAs
                    mbar_wait(smem_u32(&bars->dz_empty), (dc & 1) ^ 1);
                    dz_free = true;
                }
                uint8_t* rowp = sDz + (g >> 1) * kABlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int chunk = ((g & 1) * 4 + j) ^ (row & 7);
                    *reinterpret_cast<uint4*>(rowp + (chunk << 4)) =
                        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                // sparse part of dz: patch the (at most two) affected elements of this row in place
                const int dl = rc.lab - v0;
                const bool lab_here = dl >= 0 && dl < 32;
                const bool blank_here = blank >= v0 && blank < v0 + 32;  // warp-uniform
                auto patch = [&](int col, float corr) {
                    const int chunk = ((g & 1) * 4 + (col >> 3)) ^ (row & 7);
                    __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(rowp + (chunk << 4)) + (col & 7);
                    *e = __float2bfloat16_rn(__bfloat162float(*e) - corr);
                };
                if (lab_here) patch(dl, rc.corr_l);
                if (blank_here) patch(blank - v0, rc.corr_b);
                if (kDW && role_p == 0) {
                    float cs = warp_transpose_reduce(d, lane);   // dense part, lane == column
                    if (blank_here) {
                        const float sb = warp_sum(rc.corr_b);
                        if (lane == blank - v0) cs -= sb;
                    }
                    colsum[g] += cs;
                    if (lab_here && rc.corr_l != 0.f) atomicAdd(d_b_out + rc.lab, -rc.corr_l);
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->z_empty[zb]));
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&bars->dz_full));
            ++zc;
            ++dc;
        };

        for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            RowCtx rc;
            {
                const int m = ti.first_cell + row;
                rc.valid = m < ti.n_cells;
                const int t = rc.valid ? m / ti.U1b : 0;
                const int u = rc.valid ? m - t * ti.U1b : 0;
                rc.cell = ((size_t)ti.b * T + t) * U1 + u;
                rc.lab = -1;
                rc.c2 = 0.f; rc.gg = 0.f; rc.corr_b = 0.f; rc.corr_l = 0.f;
                if (rc.valid) {
                    const float g = __ldg(grad_cost + ti.b);
                    const float2 gm = __ldg(reinterpret_cast<const float2*>(gamma2) + rc.cell);
                    rc.c2 = -__ldg(lse + rc.cell) * kLog2e;
                    rc.gg = g * (gm.x + gm.y);
                    rc.corr_b = g * gm.x;
                    rc.corr_l = g * gm.y;
                    if (u < ti.U1b - 1)
                        rc.lab = min(max(__ldg(labels + (size_t)ti.b * (U1 - 1) + u), 0), V - 1);
                }
            }
            if (kDW) {
                epi1(role_c, rc);
            } else {
                for (int p = 0; p < NPART; ++p) {
                    const int pb = part_blocks(p);
                    for (int c = 0; c < NCH; ++c) epi1(c, rc);
                    // ---- dh_part -> dpre = dh (1 - h^2) -> bf16 (B,T,U1,J)
                    if (p == 0)
                        for (int kb = 0; kb < KB; ++kb) mbar_wait(smem_u32(&bars->a_full[kb]), tl & 1);
                    mbar_wait(smem_u32(&bars->acc2_full), ac & 1);
                    tc_fence_after();
                    for (int g = 0; g < pb * 2; ++g) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem_acc2 + lane_base + g * 32, r);
                        tmem_wait_ld();
                        const int kb = p * kPartBlocks + (g >> 1);
                        const uint8_t* hrow = sA + (size_t)kb * kABlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int chunk = ((g & 1) * 4 + j) ^ (row & 7);
                            const uint4 hv = *reinterpret_cast<const uint4*>(hrow + (chunk << 4));
                            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float h0 = __uint_as_float(hw[e] << 16);
                                const float h1 = __uint_as_float(hw[e] & 0xffff0000u);
                                const float d0 = __uint_as_float(r[j * 8 + e * 2]) * fmaf(-h0, h0, 1.f);
                                const float d1 = __uint_as_float(r[j * 8 + e * 2 + 1]) * fmaf(-h1, h1, 1.f);
                                pk[j * 4 + e] = pack_bf16x2(d0, d1);
                            }
                        }
                        if (rc.valid) {
                            uint4* dst = reinterpret_cast<uint4*>(dpre_out + rc.cell * J + kb * kBlockK + (g & 1) * 32);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars->acc2_empty));
                    ++ac;
                }
                for (int kb = 0; kb < KB; ++kb) mbar_arrive(smem_u32(&bars->a_empty[kb]));
            }
            any_tile = true;
            ++tl;
        }
        if (kDW && any_tile) {
            // ---- flush dW[c, part] (rows = vocab) and the column sums of dz
            const int pb = part_blocks(role_p);
            const int n = chunk_cols(role_c);
            mbar_wait(smem_u32(&bars->acc2_full), 0);
            tc_fence_after();
            for (int g = 0; g < pb * 2; ++g) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_acc2 + lane_base + g * 32, r);
                tmem_wait_ld();
                if (row < n) {
                    float* dst = d_w_out + (size_t)(role_c * kBwdChunk + row) * J + role_p * kPartBlocks * kBlockK + g * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
            }
            if (role_p == 0) {
#pragma unroll
                for (int g = 0; g < kBwdChunk / 32; ++g) {
                    const int v = role_c * kBwdChunk + g * 32 + lane;
                    if (g * 32 < n && v < V) atomicAdd(d_b_out + v, colsum[g]);
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== A producers =====================
        const int pw = warp - 8;
        const int c = lane & 7;
        const int rsub = lane >> 3;
        uint32_t tl = 0;
        TileInfo ti;
        for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            uint32_t eoff[8], doff[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                int row = pw * 32 + p * 4 + rsub;
                int m = min(ti.first_cell + row, ti.n_cells - 1);
                int t = m / ti.U1b, u = m - t * ti.U1b;
                eoff[p] = (uint32_t)(((size_t)ti.b * T + t) * J) + c * 8;
                doff[p] = (uint32_t)(((size_t)ti.b * U1 + u) * J) + c * 8;
            }
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(smem_u32(&bars->a_empty[kb]), (tl & 1) ^ 1);
                produce_h_block(enc, dec, eoff, doff, kb, pw, rsub, c, sA + (size_t)kb * kABlockBytes);
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&bars->a_full[kb]));
            }
            ++tl;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// d_enc_proj[b,t,j] = sum_{u <= U_b} dpre[b,t,u,j]   (0 for t >= T_b)
__global__ void reduce_over_u_kernel(const __nv_bfloat16* __restrict__ dpre, const int* __restrict__ tlen,
                                     const int* __restrict__ ulen, int T, int U1, int J,
                                     float* __restrict__ d_enc) {
    const int r = blockIdx.x;  // b*T + t
    const int b = r / T, t = r - b * T;
    const int T_b = min(max(tlen[b], 1), T), U1b = min(max(ulen[b], 0), U1 - 1) + 1;
    for (int j2 = threadIdx.x; j2 < J / 2; j2 += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        if (t < T_b) {
            const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(dpre + (size_t)r * U1 * J) + j2;
            for (int u = 0; u < U1b; ++u) {
                const float2 v = __bfloat1622float2(p[(size_t)u * (J / 2)]);
                a0 += v.x;
                a1 += v.y;
            }
        }
        reinterpret_cast<float2*>(d_enc + (size_t)r * J)[j2] = make_float2(a0, a1);
    }
}

// d_dec_proj[b,u,j] = sum_{t < T_b} dpre[b,t,u,j]   (0 for u > U_b)
__global__ void reduce_over_t_kernel(const __nv_bfloat16* __restrict__ dpre, const int* __restrict__ tlen,
                                     const int* __restrict__ ulen, int T, int U1, int J,
                                     float* __restrict__ d_dec) {
    const int r = blockIdx.x;  // b*U1 + u
    const int b = r / U1, u = r - b * U1;
    const int T_b = min(max(tlen[b], 1), T), U1b = min(max(ulen[b], 0), U1 - 1) + 1;
    for (int j2 = threadIdx.x; j2 < J / 2; j2 += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        if (u < U1b) {
            const __nv_bfloat162* p =
                reinterpret_cast<const __nv_bfloat162*>(dpre + ((size_t)b * T * U1 + u) * J) + j2;
            for (int t = 0; t < T_b; ++t) {
                const float2 v = __bfloat1622float2(p[(size_t)t * U1 * (J / 2)]);
                a0 += v.x;
                a1 += v.y;
            }
        }
        reinterpret_cast<float2*>(d_dec + (size_t)r * J)[j2] = make_float2(a0, a1);
    }
}

size_t bwd_smem_bytes(int J) {
    return (size_t)(J / kBlockK) * kABlockBytes + kDzBytes + (size_t)kBwdStages * kBwdTileBytes +
           sizeof(BwdBarriers) + 2 * kBwdChunk * sizeof(float);
}

}  // namespace

size_t joint_bf16_workspace(int op, int B, int T, int U1, int J, int V) {
    size_t w = align_up((size_t)V * J * sizeof(__nv_bfloat16), 256);
    if (op == EMO_OP_RNNT_JOINT_BWD)
        return w + align_up((size_t)B * T * U1 * J * sizeof(__nv_bfloat16), 256);
    return w;
}

int joint_bf16_launches(int op, int B, int T, int U1, int J, int V) {
    (void)B; (void)T; (void)U1; (void)J; (void)V;
    if (op == EMO_OP_RNNT_JOINT_BWD) return 5;  // weight cast, dh kernel, 2 reductions, dW kernel
    return 2;                                    // weight cast + fused joint forward
}

int joint_bwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen,
                   const float* lse, const float* gamma2, const float* grad_cost, int B, int T,
                   int U1, int J, int V, int blank, float* d_enc_proj, float* d_dec_proj,
                   float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    EMO_REQUIRE(enc_proj && dec_proj && w_out && b_out && labels && tlen && ulen && lse && gamma2 &&
                    grad_cost && d_enc_proj && d_dec_proj && d_w_out && d_b_out && ws,
                EMO_BAD_ARG, "joint_bwd(bf16): null pointer");
    int rc = check_bf16_shape(B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(ws_bytes >= joint_bf16_workspace(EMO_OP_RNNT_JOINT_BWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_bwd(bf16): workspace too small");
    EMO_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)enc_proj & 15) == 0 &&
                    ((uintptr_t)dec_proj & 15) == 0 && ((uintptr_t)w_out & 15) == 0 &&
                    ((uintptr_t)d_w_out & 15) == 0,
                EMO_BAD_ARG, "joint_bwd(bf16): pointers must be 16-byte (workspace 256-byte) aligned");
    const int KB = J / kBlockK;
    const int NCH = ceil_div(V, kBwdChunk), NPART = ceil_div(KB, kPartBlocks);
    EMO_REQUIRE(NCH * NPART <= sm_count() * 4, EMO_UNSUPPORTED_SHAPE,
                "joint_bwd(bf16): vocabulary %d too large for the dW role grid", V);
    __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* dpre = reinterpret_cast<__nv_bfloat16*>(
        (char*)ws + align_up((size_t)V * J * sizeof(__nv_bfloat16), 256));
    const size_t nw = (size_t)V * J;
    f32_to_bf16_kernel<<<ceil_div(nw, 4 * 256), 256, 0, st>>>(w_out, w_bf16, nw);
    EMO_CHECK_LAUNCH("f32_to_bf16_kernel");
    EMO_CUDA(cudaMemsetAsync(d_w_out, 0, nw * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_out, 0, (size_t)V * sizeof(float), st));

    CUtensorMap tmap;
    rc = make_tmap_bf16_2d(&tmap, w_bf16, (uint64_t)J, (uint64_t)V, kBlockK, 128);
    if (rc) return rc;
    const size_t smem = bwd_smem_bytes(J);
    EMO_REQUIRE(smem <= (size_t)kSmemLimit, EMO_UNSUPPORTED_SHAPE, "joint_bwd(bf16): shared memory");
    EMO_CUDA(cudaFuncSetAttribute(joint_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EMO_CUDA(cudaFuncSetAttribute(joint_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = B * ceil_div((size_t)T * U1, kTileM);

    // dh kernel + the two axis reductions
    joint_bwd_kernel<false><<<min(tiles, sm_count()), kThreads, smem, st>>>(
        tmap, enc_proj, dec_proj, b_out, labels, tlen, ulen, lse, gamma2, grad_cost, B, T, U1, J, V,
        blank, 1, dpre, nullptr, nullptr);
    EMO_CHECK_LAUNCH("joint_bwd_kernel<dh>");
    reduce_over_u_kernel<<<B * T, 128, 0, st>>>(dpre, tlen, ulen, T, U1, J, d_enc_proj);
    EMO_CHECK_LAUNCH("reduce_over_u_kernel");
    reduce_over_t_kernel<<<B * U1, 128, 0, st>>>(dpre, tlen, ulen, T, U1, J, d_dec_proj);
    EMO_CHECK_LAUNCH("reduce_over_t_kernel");

    // dW kernel: one (vocab chunk, J-part) role per CTA, num_splits CTAs per role
    const int roles = NCH * NPART;
    const int splits = max(1, min(sm_count() / roles, tiles));
    joint_bwd_kernel<true><<<roles * splits, kThreads, smem, st>>>(
        tmap, enc_proj, dec_proj, b_out, labels, tlen, ulen, lse, gamma2, grad_cost, B, T, U1, J, V,
        blank, splits, nullptr, d_w_out, d_b_out);
    EMO_CHECK_LAUNCH("joint_bwd_kernel<dW>");
    return EMO_OK;
}

}  // namespace emo

// EMO_PREC_BF16 backward of the fused joint (tcgen05 / TMEM / TMA).
//
// The dense gradient of the logits
//     dz[cell,v] = g_b * (gamma * softmax(z)[v] - gamma_blank 1[v=blank] - gamma_label 1[v=label])
// is recomputed tile by tile from (h, w_out, lse) and consumed straight from shared memory by the
// MMAs that need it; it never reaches global memory.  h = tanh(enc+dec) comes from the bf16 h
// cache the forward kernel wrote (TMA loads, no tanh / fp32 gathers here).  Two kernels:
//
//   dh kernel   cell-stationary CTA PAIR (cluster of 2, tcgen05 cta_group::2, 256 cells per pair
//               tile).  Per J-part (<= 256 hidden units) loop over 128-wide vocab chunks:
//                  z = h W_c^T (M=256) -> dz (bf16, smem) -> dh_part += dz W_c[:, part]
//               then dpre = dh (1 - h^2) is written as bf16 (B,T,U1,J) for the two axis reductions
//               (d_enc_proj = sum_u, d_dec_proj = sum_t) done by small kernels.  The pair halves
//               the w_out traffic per cell and the shared-memory operand traffic per MMA.
//   dW kernel   one (vocab chunk, J-part) role per CTA, persistent over 128-cell tiles:
//                  z = h W_c^T -> dz -> dW[c, part] += dz^T h[:, part]
//               (accumulates in TMEM over ALL tiles of the CTA, flushed once with
//               red.global.add.v4.f32); the same CTAs produce d_b_out = column sums of dz.
//
// Operand layouts: h and dz tiles are stored once, rows = cells, 128-byte rows of 64 bf16, 128B
// swizzle.  Read as K-major they feed z = h W^T and dh = dz W; read as MN-major (same bytes,
// different descriptor) they feed dW = dz^T h.  w_out tiles arrive by TMA and are K-major B for z
// and MN-major B for dh.
// Both kernels: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue
// (two warps per TMEM lane quadrant, each owning half of the columns of a chunk).
// TMEM: columns [0,256) two z buffers of 128, columns [256,512) the dh / dW accumulator.
#include "joint_tc.cuh"

namespace emo {
namespace {

constexpr int kBwdChunk = 128;                 // vocab columns per z chunk
constexpr int kSlotBytes = 16384;              // one ring slot = one [128 x 64] bf16 tile
constexpr int kDhSlots = 4;                    // dh kernel: w_out ring
constexpr int kDwSlots = 4;                    // dW kernel: w_out ring
constexpr int kDzBytes = 2 * kABlockBytes;     // [128 cells x 128 v] bf16 = 32 KiB
constexpr int kPartBlocks = 4;                 // J-part = up to 4 K blocks = 256 hidden units
constexpr int kEpiThreads = 256;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024, v1, SW128

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

__device__ __forceinline__ void load_row_ctx(RowCtx& rc, const TileInfo& ti, int row, int T, int U1,
                                             int V, const int* __restrict__ labels,
                                             const float* __restrict__ lse,
                                             const float* __restrict__ gamma2,
                                             const float* __restrict__ grad_cost) {
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

// One z chunk (TMEM, this thread's row) -> dz (bf16) for the two 32-column groups this thread owns
// (column half `hf` of the 128-wide chunk), written into the shared dz tile [128 cells x 128 v]
// (two K-major SW128 atoms).  kColSum: also accumulate the per-column sums (d_b_out).
template <bool kColSum>
__device__ __forceinline__ void dz_from_z(uint32_t taddr, const float* __restrict__ bias, int n, int v0c,
                                          int hf, int row, int lane, const RowCtx& rc, int blank,
                                          uint8_t* sDz, uint32_t dz_empty_bar, uint32_t dz_empty_parity,
                                          float (&colsum)[2], float* __restrict__ d_b_out) {
    const int g0 = hf * 2;
    if (g0 * 32 >= n) return;
    const bool two = (g0 + 1) * 32 < n;
    uint32_t r0[32], r1[32];
    tmem_ld_32x32b_x32(taddr + g0 * 32, r0);
    if (two) tmem_ld_32x32b_x32(taddr + g0 * 32 + 32, r1);
    tmem_wait_ld();
    uint8_t* rowp = sDz + hf * kABlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
    bool dz_free = false;
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        if (gi == 1 && !two) break;
        const uint32_t(&r)[32] = gi == 0 ? r0 : r1;
        const int v0 = v0c + (g0 + gi) * 32;
        const float* bs = bias + (g0 + gi) * 32;
        float d[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bs + i);
            d[i + 0] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 0]) + bv.x, kLog2e, rc.c2));
            d[i + 1] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 1]) + bv.y, kLog2e, rc.c2));
            d[i + 2] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 2]) + bv.z, kLog2e, rc.c2));
            d[i + 3] = rc.gg * ex2_approx(fmaf(__uint_as_float(r[i + 3]) + bv.w, kLog2e, rc.c2));
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(d[2 * i], d[2 * i + 1]);
        if (!dz_free) {  // the previous dz tile must have been consumed by its MMAs
            mbar_wait(dz_empty_bar, dz_empty_parity);
            dz_free = true;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int chunk = (gi * 4 + j) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + (chunk << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        // sparse part of dz: patch the (at most two) affected elements of this row in place
        const int dl = rc.lab - v0;
        const bool lab_here = dl >= 0 && dl < 32;
        const bool blank_here = blank >= v0 && blank < v0 + 32;  // warp-uniform
        auto patch = [&](int col, float corr) {
            const int chunk = (gi * 4 + (col >> 3)) ^ (row & 7);
            __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(rowp + (chunk << 4)) + (col & 7);
            *e = __float2bfloat16_rn(__bfloat162float(*e) - corr);
        };
        if (lab_here) patch(dl, rc.corr_l);
        if (blank_here) patch(blank - v0, rc.corr_b);
        if (kColSum) {
            float cs = warp_transpose_reduce(d, lane);   // dense part, lane == column
            if (blank_here) {
                const float sb = warp_sum(rc.corr_b);
                if (lane == blank - v0) cs -= sb;
            }
            colsum[gi] += cs;
            if (lab_here && rc.corr_l != 0.f) atomicAdd(d_b_out + rc.lab, -rc.corr_l);
        }
    }
}

// dpre = dacc (1 - h^2) for one 32-column group of this thread's row; h from the resident h tile
__device__ __forceinline__ void dpre_group(const uint32_t (&r)[32], const uint8_t* hrow, int half32, int row,
                                           uint32_t (&pk)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int chunk = (half32 * 4 + j) ^ (row & 7);
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
}

// =================================================================================================
// dh kernel (CTA pair).  Barrier topology: *_full barriers fed by TMA and the barriers the MMA issuer
// waits on (z_empty, dz_full, acc_empty) live in the LEADER; barriers signalled by tcgen05.commit
// (w_empty, z_full, dz_empty, acc_full, h_empty) are multicast to BOTH CTAs.
struct __align__(16) DhBarriers {
    uint64_t w_full[kDhSlots], w_empty[kDhSlots];
    uint64_t h_full[kMaxKBlocks];
    uint64_t h_empty[kMaxKBlocks];   // per K block: blocks outside the last J-part are free before the last epilogue
    uint64_t z_full[2], z_empty[2];
    uint64_t dz_full, dz_empty;
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__global__ void __launch_bounds__(kThreads, 1)
joint_dh_kernel(const __grid_constant__ CUtensorMap tmap_wz,   // w_out bf16, box [64 j x 64 v]
                const __grid_constant__ CUtensorMap tmap_wd,   // w_out bf16, box [64 j x 128 v]
                const __grid_constant__ CUtensorMap tmap_h,    // h cache, box [64 j x 128 cells]
                const float* __restrict__ b_out, const int* __restrict__ labels,
                const int* __restrict__ tlen, const int* __restrict__ ulen,
                const float* __restrict__ lse, const float* __restrict__ gamma2,
                const float* __restrict__ grad_cost, int B, int T, int U1, int J, int V, int blank,
                __nv_bfloat16* __restrict__ dpre_out) {   // (B,T,U1,J)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = J / kBlockK;
    const int NCH = (V + kBwdChunk - 1) / kBwdChunk;
    const int NPART = (KB + kPartBlocks - 1) / kPartBlocks;
    const int ZS = (KB + 1) / 2;  // ring slots per z chunk (two K blocks of [64 v x 64 j] per slot)
    uint8_t* sH = smem;
    uint8_t* sDz = sH + (size_t)KB * kABlockBytes;
    uint8_t* sW = sDz + kDzBytes;
    DhBarriers* bars = reinterpret_cast<DhBarriers*>(sW + (size_t)kDhSlots * kSlotBytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);  // [2][kBwdChunk]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_per_utt = (T * U1 + 2 * kTileM - 1) / (2 * kTileM);
    const int total_tiles = B * tiles_per_utt;
    const int tile0 = blockIdx.x / 2, tile_stride = gridDim.x / 2;
    const int tpu = tiles128_per_utt(T, U1);
    auto chunk_cols = [&](int c) { return min(kBwdChunk, V - c * kBwdChunk); };
    auto part_cols = [&](int p) { return min(kPartBlocks * kBlockK, J - p * kPartBlocks * kBlockK); };

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kDhSlots; ++i) {
            mbar_init(smem_u32(&bars->w_full[i]), 2);
            mbar_init(smem_u32(&bars->w_empty[i]), 1);
        }
        for (int i = 0; i < kMaxKBlocks; ++i) mbar_init(smem_u32(&bars->h_full[i]), 2);
        // a block is free once the last z MMA has read it (commit) and, for the blocks of the LAST J-part,
        // once the dpre epilogue of that part has read h from it as well (256 threads)
        for (int i = 0; i < kMaxKBlocks; ++i)
            mbar_init(smem_u32(&bars->h_empty[i]), i >= (NPART - 1) * kPartBlocks ? 1 + kEpiThreads : 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->z_empty[i]), 2 * kEpiWarps);   // one arrive per epilogue warp
        }
        mbar_init(smem_u32(&bars->dz_full), 2 * kEpiWarps);
        mbar_init(smem_u32(&bars->dz_empty), 1);
        mbar_init(smem_u32(&bars->acc_full), 1);
        mbar_init(smem_u32(&bars->acc_empty), 2 * kEpiWarps);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_wz);
        tma_prefetch_desc(&tmap_wd);
        tma_prefetch_desc(&tmap_h);
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t tmem_acc2 = tmem_base + 2 * kBwdChunk;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; same order as the MMA issuer consumes) =====
        if (lane == 0) {
            uint32_t slot = 0, sphase = 0, tl = 0;
            uint32_t full = 0;
            auto acquire = [&](uint32_t bytes) -> uint32_t {
                mbar_wait(smem_u32(&bars->w_empty[slot]), sphase ^ 1);
                full = smem_u32(&bars->w_full[slot]);
                mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), bytes);
                const uint32_t dst = smem_u32(sW + (size_t)slot * kSlotBytes);
                if (++slot == kDhSlots) { slot = 0; sphase ^= 1; }
                return dst;
            };
            auto zloads = [&](int c) {
                const int y = c * kBwdChunk + (int)rank * (chunk_cols(c) >> 1);
                for (int s = 0; s < ZS; ++s) {
                    const int nkb = min(2, KB - 2 * s);
                    const uint32_t dst = acquire(nkb * (kSlotBytes / 2));
                    for (int i = 0; i < nkb; ++i)
                        tma_load_2d_pair(dst + i * (kSlotBytes / 2), &tmap_wz, (2 * s + i) * kBlockK, y, full);
                }
            };
            auto dhloads = [&](int p, int c) {
                const int half = part_cols(p) >> 1;
                for (int t = 0; t < half / kBlockK; ++t) {
                    const uint32_t dst = acquire(kSlotBytes);
                    tma_load_2d_pair(dst, &tmap_wd, p * kPartBlocks * kBlockK + (int)rank * half + t * kBlockK,
                                     c * kBwdChunk, full);
                }
            };
            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<2>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
                for (int p = 0; p < NPART; ++p) {
                    zloads(0);
                    for (int c = 0; c < NCH; ++c) {
                        if (c + 1 < NCH) zloads(c + 1);
                        dhloads(p, c);
                    }
                }
                ++tl;
            }
        }
    } else if (warp == 3) {
        // ===================== TMA: the h tile, block by block as the previous tile lets go of it (own thread,
        // so a block still held by the last epilogue never delays the w_out stream) =====================
        if (lane == 0) {
            uint32_t tl = 0;
            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<2>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
                const int row0 = (ti.b * tpu + ti.first_cell / kTileM) * kTileM;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(smem_u32(&bars->h_empty[kb]), (tl & 1) ^ 1);
                    const uint32_t hf = smem_u32(&bars->h_full[kb]);
                    mbar_arrive_expect_tx_cluster(mapa_shared(hf, 0), kABlockBytes);
                    tma_load_2d_pair(smem_u32(sH + (size_t)kb * kABlockBytes), &tmap_h, kb * kBlockK, row0, hf);
                }
                ++tl;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; one elected lane issues) =====================
        if (leader) {
            uint32_t slot = 0, sphase = 0, zc = 0, dc = 0, ac = 0, tl = 0;
            const uint32_t h_lo0 = desc_lo(smem_u32(sH), 16);
            const uint32_t w_lo0 = desc_lo(smem_u32(sW), 16);
            const uint32_t dz_lo0 = desc_lo(smem_u32(sDz), 16);
            const uint32_t w_mn_lo0 = desc_lo(smem_u32(sW), kSlotBytes);
            auto advance = [&]() { if (++slot == kDhSlots) { slot = 0; sphase ^= 1; } };

            // z[256 x n] = h W_c^T into z buffer zc&1
            auto z_mma = [&](int c, bool wait_h, bool release_h) {
                const uint32_t zb = zc & 1;
                mbar_wait(smem_u32(&bars->z_empty[zb]), ((zc >> 1) & 1) ^ 1);
                const uint32_t idesc = umma_idesc_bf16(2 * kTileM, chunk_cols(c));
                const uint32_t d_tmem = tmem_base + zb * kBwdChunk;
                for (int s = 0; s < ZS; ++s) {
                    const int nkb = min(2, KB - 2 * s);
                    if (wait_h)
                        for (int i = 0; i < nkb; ++i) mbar_wait(smem_u32(&bars->h_full[2 * s + i]), tl & 1);
                    mbar_wait(smem_u32(&bars->w_full[slot]), sphase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        for (int i = 0; i < nkb; ++i) {
                            const uint32_t a_lo = h_lo0 + (2 * s + i) * (kABlockBytes >> 4);
                            const uint32_t b_lo = w_lo0 + slot * (kSlotBytes >> 4) + i * (kSlotBytes >> 5);
#pragma unroll
                            for (int k16 = 0; k16 < kBlockK / 16; ++k16)
                                umma_bf16_pair(d_tmem, mk_desc(a_lo + 2 * k16), mk_desc(b_lo + 2 * k16), idesc,
                                               (s | i | k16) != 0);
                        }
                        umma_commit_pair(smem_u32(&bars->w_empty[slot]));
                        if (s == ZS - 1) {
                            umma_commit_pair(smem_u32(&bars->z_full[zb]));
                            if (release_h)
                                for (int kb = 0; kb < KB; ++kb) umma_commit_pair(smem_u32(&bars->h_empty[kb]));
                        }
                    }
                    __syncwarp();
                    advance();
                }
                ++zc;
            };
            // dh_part[256 x part] += dz W_c[:, part]; one N=128 MMA group per w_out tile (64 j per CTA)
            auto dh_mma = [&](int p, int c) {
                const int n = chunk_cols(c);
                const int nt = (part_cols(p) >> 1) / kBlockK;
                if (c == 0) mbar_wait(smem_u32(&bars->acc_empty), (ac & 1) ^ 1);
                mbar_wait(smem_u32(&bars->dz_full), dc & 1);
                const uint32_t idesc = umma_idesc_bf16(2 * kTileM, 2 * kBlockK, 0, 1);
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(smem_u32(&bars->w_full[slot]), sphase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t b_lo = w_mn_lo0 + slot * (kSlotBytes >> 4);
                        for (int kk = 0; kk < n / 16; ++kk)
                            umma_bf16_pair(tmem_acc2 + t * 2 * kBlockK,
                                           mk_desc(dz_lo0 + (kk >> 2) * (kABlockBytes >> 4) + (kk & 3) * 2),
                                           mk_desc(b_lo + kk * (2048 >> 4)), idesc, (c > 0 || kk > 0) ? 1u : 0u);
                        umma_commit_pair(smem_u32(&bars->w_empty[slot]));
                        if (t == nt - 1) {
                            umma_commit_pair(smem_u32(&bars->dz_empty));
                            if (c == NCH - 1) umma_commit_pair(smem_u32(&bars->acc_full));
                        }
                    }
                    __syncwarp();
                    advance();
                }
                ++dc;
            };

            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<2>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
                for (int p = 0; p < NPART; ++p) {
                    const bool lastp = p == NPART - 1;
                    z_mma(0, p == 0, lastp && NCH == 1);
                    for (int c = 0; c < NCH; ++c) {
                        if (c + 1 < NCH) z_mma(c + 1, false, lastp && c + 1 == NCH - 1);
                        dh_mma(p, c);
                    }
                    ++ac;
                }
                ++tl;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        const int e = threadIdx.x - 128;
        const int q = warp & 3, hf = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t zc = 0, dc = 0, ac = 0;
        float colsum[2] = {0.f, 0.f};
        const uint32_t z_empty_addr[2] = {mapa_shared(smem_u32(&bars->z_empty[0]), 0),
                                          mapa_shared(smem_u32(&bars->z_empty[1]), 0)};
        const uint32_t dz_full_addr = mapa_shared(smem_u32(&bars->dz_full), 0);
        const uint32_t acc_empty_addr = mapa_shared(smem_u32(&bars->acc_empty), 0);
        TileInfo ti;
        float nb = e < chunk_cols(0) ? __ldg(b_out + e) : 0.f;   // bias of the next chunk, one chunk ahead
        for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
            if (!tile_info<2>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
            RowCtx rc;
            load_row_ctx(rc, ti, row, T, U1, V, labels, lse, gamma2, grad_cost);
            for (int p = 0; p < NPART; ++p) {
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t zb = zc & 1;
                    const int n = chunk_cols(c);
                    float* bias = s_bias + zb * kBwdChunk;
                    if (e < kBwdChunk) bias[e] = nb;
                    {
                        const int cn = c + 1 == NCH ? 0 : c + 1;
                        nb = e < chunk_cols(cn) ? __ldg(b_out + cn * kBwdChunk + e) : 0.f;
                    }
                    named_bar_sync(1, kEpiThreads);
                    mbar_wait(smem_u32(&bars->z_full[zb]), (zc >> 1) & 1);
                    tc_fence_after();
                    dz_from_z<false>(tmem_base + lane_base + zb * kBwdChunk, bias, n, c * kBwdChunk, hf, row, lane,
                                     rc, blank, sDz, smem_u32(&bars->dz_empty), (dc & 1) ^ 1, colsum, nullptr);
                    // every lane orders its own TMEM reads / shared-memory writes, then ONE lane per warp
                    // arrives (512 single-word arrivals per chunk would serialise on the barrier)
                    tc_fence_before();
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_cluster(z_empty_addr[zb]);
                        mbar_arrive_cluster(dz_full_addr);
                    }
                    ++zc;
                    ++dc;
                }
                // ---- dh_part -> dpre = dh (1 - h^2) -> bf16 (B,T,U1,J).  Accumulator columns of the
                // part: [t*128 + r*64 + jj] <-> j = p*256 + r*half + t*64 + jj  (r = CTA that staged it)
                const int half = part_cols(p) >> 1;
                const int G = part_cols(p) >> 5;  // 32-column groups
                mbar_wait(smem_u32(&bars->acc_full), ac & 1);
                tc_fence_after();
                for (int gi = hf * (G >> 1); gi < (hf + 1) * (G >> 1); ++gi) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tmem_acc2 + lane_base + gi * 32, r);
                    tmem_wait_ld();
                    const int j0 = p * kPartBlocks * kBlockK + ((gi >> 1) & 1) * half + (gi >> 2) * kBlockK + (gi & 1) * 32;
                    const int kb = j0 / kBlockK;
                    const uint8_t* hrow = sH + (size_t)kb * kABlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
                    uint32_t pk[16];
                    dpre_group(r, hrow, gi & 1, row, pk);
                    if (rc.valid) {
                        uint4* dst = reinterpret_cast<uint4*>(dpre_out + rc.cell * J + j0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty_addr);
                ++ac;
            }
            for (int kb = (NPART - 1) * kPartBlocks; kb < KB; ++kb) mbar_arrive(smem_u32(&bars->h_empty[kb]));
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// =================================================================================================
// dW kernel (single CTA per role, cta_group::1).
struct __align__(16) DwBarriers {
    uint64_t w_full[kDwSlots], w_empty[kDwSlots];
    uint64_t h_full[kMaxKBlocks], h_empty[kMaxKBlocks];
    uint64_t z_full[2], z_empty[2];
    uint64_t dz_full, dz_empty;
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__global__ void __launch_bounds__(kThreads, 1)
joint_dw_kernel(const __grid_constant__ CUtensorMap tmap_w,    // w_out bf16, box [64 j x 128 v]
                const __grid_constant__ CUtensorMap tmap_h,    // h cache, box [64 j x 128 cells]
                const float* __restrict__ b_out, const int* __restrict__ labels,
                const int* __restrict__ tlen, const int* __restrict__ ulen,
                const float* __restrict__ lse, const float* __restrict__ gamma2,
                const float* __restrict__ grad_cost, int B, int T, int U1, int J, int V, int blank,
                int num_splits,
                float* __restrict__ d_w_out,    // (V,J), pre-zeroed
                float* __restrict__ d_b_out) {  // (V), pre-zeroed
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = J / kBlockK;
    const int NCH = (V + kBwdChunk - 1) / kBwdChunk;
    const int NPART = (KB + kPartBlocks - 1) / kPartBlocks;
    uint8_t* sH = smem;                                   // h tile: KB blocks [128 cells x 64 j]
    uint8_t* sDz = sH + (size_t)KB * kABlockBytes;
    uint8_t* sW = sDz + kDzBytes;                         // ring of W_c blocks [128 v x 64 j]
    DwBarriers* bars = reinterpret_cast<DwBarriers*>(sW + (size_t)kDwSlots * kSlotBytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);   // [kBwdChunk]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_utt = (T * U1 + kTileM - 1) / kTileM;
    const int total_tiles = B * tiles_per_utt;
    const int tpu = tiles128_per_utt(T, U1);
    // role = (vocab chunk, J-part); `num_splits` CTAs per role share the cell tiles
    const int role = (int)blockIdx.x % (NCH * NPART);
    const int role_c = role / NPART, role_p = role % NPART;
    const int tile0 = (int)blockIdx.x / (NCH * NPART);
    const int n = min(kBwdChunk, V - role_c * kBwdChunk);            // vocab rows of the role
    const int pb = min(kPartBlocks, KB - role_p * kPartBlocks);      // K blocks of the role's J-part
    const int kb_part0 = role_p * kPartBlocks;
    // K order of the z MMAs: the blocks outside the part first (their h slots are recycled early)
    auto korder = [&](int i) { return i < KB - pb ? (i < kb_part0 ? i : i + pb) : kb_part0 + (i - (KB - pb)); };

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kDwSlots; ++i) {
            mbar_init(smem_u32(&bars->w_full[i]), 1);
            mbar_init(smem_u32(&bars->w_empty[i]), 1);
        }
        for (int i = 0; i < kMaxKBlocks; ++i) {
            mbar_init(smem_u32(&bars->h_full[i]), 1);
            mbar_init(smem_u32(&bars->h_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->z_empty[i]), kEpiThreads);
        }
        mbar_init(smem_u32(&bars->dz_full), kEpiThreads);
        mbar_init(smem_u32(&bars->dz_empty), 1);
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_h);
    }
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
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t tl = 0, slot = 0, sphase = 0;
            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += num_splits) {
                if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
                const int row0 = (ti.b * tpu + ti.first_cell / kTileM) * kTileM;
                for (int i = 0; i < KB; ++i) {
                    const int kb = korder(i);
                    mbar_wait(smem_u32(&bars->h_empty[kb]), (tl & 1) ^ 1);
                    const uint32_t hf = smem_u32(&bars->h_full[kb]);
                    mbar_arrive_expect_tx(hf, kABlockBytes);
                    tma_load_2d(smem_u32(sH + (size_t)kb * kABlockBytes), &tmap_h, kb * kBlockK, row0, hf);
                    mbar_wait(smem_u32(&bars->w_empty[slot]), sphase ^ 1);
                    const uint32_t wf = smem_u32(&bars->w_full[slot]);
                    mbar_arrive_expect_tx(wf, kSlotBytes);
                    tma_load_2d(smem_u32(sW + (size_t)slot * kSlotBytes), &tmap_w, kb * kBlockK, role_c * kBwdChunk, wf);
                    if (++slot == kDwSlots) { slot = 0; sphase ^= 1; }
                }
                ++tl;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t zc = 0, dc = 0, slot = 0, sphase = 0;
        const uint32_t h_lo0 = desc_lo(smem_u32(sH), 16);
        const uint32_t w_lo0 = desc_lo(smem_u32(sW), 16);
        const uint32_t h_mn_lo0 = desc_lo(smem_u32(sH), kABlockBytes);
        const uint32_t dz_mn_lo0 = desc_lo(smem_u32(sDz), kABlockBytes);
        const uint32_t idesc_z = umma_idesc_bf16(kTileM, n);
        const uint32_t idesc_dw = umma_idesc_bf16(kTileM, pb * kBlockK, 1, 1);
        // z MMAs over K blocks korder(i0..i1) of the tile with parity `par`; z buffer zc&1
        auto z_part = [&](int i0, int i1, uint32_t par, bool first, bool last) {
            const uint32_t zb = zc & 1;
            if (first) mbar_wait(smem_u32(&bars->z_empty[zb]), ((zc >> 1) & 1) ^ 1);
            const uint32_t d_tmem = tmem_base + zb * kBwdChunk;
            for (int i = i0; i < i1; ++i) {
                const int kb = korder(i);
                mbar_wait(smem_u32(&bars->h_full[kb]), par);
                mbar_wait(smem_u32(&bars->w_full[slot]), sphase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_lo = h_lo0 + kb * (kABlockBytes >> 4);
                    const uint32_t b_lo = w_lo0 + slot * (kSlotBytes >> 4);
#pragma unroll
                    for (int k16 = 0; k16 < kBlockK / 16; ++k16)
                        umma_bf16(d_tmem, mk_desc(a_lo + 2 * k16), mk_desc(b_lo + 2 * k16), idesc_z,
                                  (i | k16) != 0);
                    umma_commit(smem_u32(&bars->w_empty[slot]));
                    if (i < KB - pb) umma_commit(smem_u32(&bars->h_empty[kb]));  // outside the part: free now
                    if (last && i == i1 - 1) umma_commit(smem_u32(&bars->z_full[zb]));
                }
                __syncwarp();
                if (++slot == kDwSlots) { slot = 0; sphase ^= 1; }
            }
            if (last) ++zc;
        };
        // dW[c, part] += dz^T h[:, part]   (M = 128 vocab rows, N = 64 pb, K = 128 cells)
        auto dw_mma = [&](bool accumulate) {
            mbar_wait(smem_u32(&bars->dz_full), dc & 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t hb_lo = h_mn_lo0 + kb_part0 * (kABlockBytes >> 4);
#pragma unroll
                for (int kk = 0; kk < kTileM / 16; ++kk)
                    umma_bf16(tmem_acc2, mk_desc(dz_mn_lo0 + kk * (2048 >> 4)), mk_desc(hb_lo + kk * (2048 >> 4)),
                              idesc_dw, (accumulate || kk) ? 1u : 0u);
                umma_commit(smem_u32(&bars->dz_empty));
                for (int jb = 0; jb < pb; ++jb) umma_commit(smem_u32(&bars->h_empty[kb_part0 + jb]));
            }
            __syncwarp();
            ++dc;
        };

        // software pipeline over the CTA's valid tiles: zA(next) is issued before dW(cur) so the
        // tensor pipe has work while the epilogue turns z(cur) into dz(cur)
        TileInfo ti;
        uint32_t tl = 0;
        bool have_cur = false;
        for (int tile = tile0; tile < total_tiles; tile += num_splits) {
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            // `tile` is the next valid tile (index tl); the current one (tl-1) still owes its dW
            z_part(0, KB - pb, tl & 1, true, false);
            if (have_cur) dw_mma(tl > 1);
            z_part(KB - pb, KB, tl & 1, KB - pb == 0, true);
            have_cur = true;
            ++tl;
        }
        if (have_cur) {
            dw_mma(tl > 1);
            if (elect_one_sync()) umma_commit(smem_u32(&bars->acc_full));
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        const int e = threadIdx.x - 128;
        const int q = warp & 3, hf = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t zc = 0, dc = 0;
        float colsum[2] = {0.f, 0.f};
        bool any_tile = false;
        if (e < n) s_bias[e] = __ldg(b_out + role_c * kBwdChunk + e);
        named_bar_sync(1, kEpiThreads);
        TileInfo ti;
        for (int tile = tile0; tile < total_tiles; tile += num_splits) {
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            RowCtx rc;
            load_row_ctx(rc, ti, row, T, U1, V, labels, lse, gamma2, grad_cost);
            const uint32_t zb = zc & 1;
            mbar_wait(smem_u32(&bars->z_full[zb]), (zc >> 1) & 1);
            tc_fence_after();
            if (role_p == 0)
                dz_from_z<true>(tmem_base + lane_base + zb * kBwdChunk, s_bias, n, role_c * kBwdChunk, hf, row, lane,
                                rc, blank, sDz, smem_u32(&bars->dz_empty), (dc & 1) ^ 1, colsum, d_b_out);
            else
                dz_from_z<false>(tmem_base + lane_base + zb * kBwdChunk, s_bias, n, role_c * kBwdChunk, hf, row, lane,
                                 rc, blank, sDz, smem_u32(&bars->dz_empty), (dc & 1) ^ 1, colsum, d_b_out);
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->z_empty[zb]));
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&bars->dz_full));
            ++zc;
            ++dc;
            any_tile = true;
        }
        if (any_tile) {
            // ---- flush dW[c, part] (rows = vocab) and the column sums of dz
            mbar_wait(smem_u32(&bars->acc_full), 0);
            tc_fence_after();
            const int G = pb * 2;  // 32-column groups of the part
            for (int g = hf * (G >> 1); g < (hf + 1) * (G >> 1); ++g) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_acc2 + lane_base + g * 32, r);
                tmem_wait_ld();
                if (row < n) {
                    float* dst = d_w_out + (size_t)(role_c * kBwdChunk + row) * J + kb_part0 * kBlockK + g * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
            }
            if (role_p == 0) {
#pragma unroll
                for (int gi = 0; gi < 2; ++gi) {
                    const int col = (hf * 2 + gi) * 32 + lane;
                    if (col < n) atomicAdd(d_b_out + role_c * kBwdChunk + col, colsum[gi]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =================================================================================================
// dW kernel, transposed CTA-pair version ("dWt").  Role of a pair = (256-wide vocab range, 256-wide J
// group); both CTAs work on the SAME 128-cell tile and the epilogue thread is a VOCAB row:
//   z^T [256 v x 128 cells] = W_range (A: this CTA's 128 vocab rows, RESIDENT in shared memory)
//                             . h^T   (B: the h tile, split by cells across the pair -> 64 cells x J per
//                                         CTA, streamed through a small ring)
//   dz^T[v, cell] (bf16)  rows = vocab, 128 cells contiguous  ->  K-major B operand of
//   dW^T [256 j x 256 v] += h^T (A: MN-major view of this CTA's [128 cells x 128 j] block of the h tile)
//                           . dz (B: dz^T rows, split by vocab across the pair)        K = 128 cells
// accumulated in TMEM over ALL tiles of the pair and flushed once.  Per tile a CTA pulls 96 KiB of h
// from L2 and nothing else (w_out never streams), and the h block used by dW^T is loaded separately
// from the ring that feeds z^T, so no operand is held across the epilogue.
constexpr int kHzSlots = 4;
constexpr int kHzBytes = 64 * kBlockK * 2;   // [64 cells x 64 j] bf16 = 8 KiB

struct __align__(16) DwtBarriers {
    uint64_t w_full;
    uint64_t hz_full[kHzSlots], hz_empty[kHzSlots];
    uint64_t hd_full, hd_empty;
    uint64_t z_full[2], z_empty[2];
    uint64_t dz_full, dz_empty;
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__global__ void __launch_bounds__(kThreads, 1)
joint_dwt_kernel(const __grid_constant__ CUtensorMap tmap_w,    // w_out bf16, box [64 j x 128 v]
                 const __grid_constant__ CUtensorMap tmap_hz,   // h cache, box [64 j x 64 cells]
                 const __grid_constant__ CUtensorMap tmap_hd,   // h cache, box [64 j x 128 cells]
                 const float* __restrict__ b_out, const int* __restrict__ labels,
                 const int* __restrict__ tlen, const int* __restrict__ ulen,
                 const float* __restrict__ lse, const float* __restrict__ gamma2,
                 const float* __restrict__ grad_cost, int B, int T, int U1, int J, int V, int blank,
                 int num_splits,
                 float* __restrict__ d_w_out,    // (V,J), pre-zeroed
                 float* __restrict__ d_b_out) {  // (V), pre-zeroed
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = J / kBlockK;
    const int G = (J + 255) / 256;                        // J groups
    uint8_t* sW = smem;                                   // resident: KB blocks [128 v x 64 j]
    uint8_t* sHz = sW + (size_t)KB * kABlockBytes;        // ring of [64 cells x 64 j] blocks (z^T B operand)
    uint8_t* sHd = sHz + (size_t)kHzSlots * kHzBytes;     // [128 cells x 128 j] (dW^T A operand), 2 x 16 KiB
    uint8_t* sDz = sHd + 2 * kABlockBytes;                // dz^T [128 v x 128 cells], 2 K-atoms x 16 KiB
    DwtBarriers* bars = reinterpret_cast<DwtBarriers*>(sDz + kDzBytes);
    float* s_c2 = reinterpret_cast<float*>(bars + 1);     // per-cell scalars of the current tile
    float* s_gg = s_c2 + kTileM;
    float* s_cb = s_gg + kTileM;
    float* s_cl = s_cb + kTileM;
    int* s_lab = reinterpret_cast<int*>(s_cl + kTileM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_per_utt = (T * U1 + kTileM - 1) / kTileM;
    const int total_tiles = B * tiles_per_utt;
    const int tpu = tiles128_per_utt(T, U1);
    const int pair = blockIdx.x >> 1;
    const int roles = G * ((V + 255) / 256);
    const int role = pair % roles, tile0 = pair / roles;
    const int g = role % G, range = role / G;
    const int v0 = range * 256, v0_r = v0 + (int)rank * kTileM;       // first vocab row of the pair / this CTA
    const bool full_group = J - g * 256 >= 256;                      // 128-wide last group: both CTAs duplicate it
    const int j0_r = g * 256 + (full_group ? (int)rank * kTileM : 0);
    const bool flush_dw = full_group || leader;
    // The pairs of one split walk the same tile list; each role starts at a different point of it, so the
    // 2 * roles CTAs do not all pull the same h tile out of the same L2 lines at the same moment.
    const int n_seq = tile0 < total_tiles ? (total_tiles - tile0 + num_splits - 1) / num_splits : 0;
    const int seq_off = n_seq > 0 ? (role * 5) % n_seq : 0;
    auto tile_at = [&](int sk) { int i = sk + seq_off; if (i >= n_seq) i -= n_seq; return tile0 + i * num_splits; };

    if (warp == 1 && lane == 0) {
        mbar_init(smem_u32(&bars->w_full), 2);
        for (int i = 0; i < kHzSlots; ++i) {
            mbar_init(smem_u32(&bars->hz_full[i]), 2);
            mbar_init(smem_u32(&bars->hz_empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->hd_full), 2);
        mbar_init(smem_u32(&bars->hd_empty), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->z_empty[i]), 2 * kEpiWarps);   // one arrive per epilogue warp
        }
        mbar_init(smem_u32(&bars->dz_full), 2 * kEpiWarps);
        mbar_init(smem_u32(&bars->dz_empty), 1);
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_hz);
        tma_prefetch_desc(&tmap_hd);
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t tmem_acc2 = tmem_base + 2 * kBwdChunk;

    if (warp == 0) {
        // ===================== TMA: resident W, then the h ring feeding z^T =====================
        if (lane == 0) {
            const uint32_t wf = smem_u32(&bars->w_full);
            mbar_arrive_expect_tx_cluster(mapa_shared(wf, 0), KB * kABlockBytes);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d_pair(smem_u32(sW + (size_t)kb * kABlockBytes), &tmap_w, kb * kBlockK, v0_r, wf);
            uint32_t slot = 0, sphase = 0;
            TileInfo ti;
            for (int sk = 0; sk < n_seq; ++sk) {
            const int tile = tile_at(sk);
                if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
                const int row0 = (ti.b * tpu + ti.first_cell / kTileM) * kTileM + (int)rank * 64;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(smem_u32(&bars->hz_empty[slot]), sphase ^ 1);
                    const uint32_t full = smem_u32(&bars->hz_full[slot]);
                    mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), kHzBytes);
                    tma_load_2d_pair(smem_u32(sHz + (size_t)slot * kHzBytes), &tmap_hz, kb * kBlockK, row0, full);
                    if (++slot == kHzSlots) { slot = 0; sphase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== TMA: the [128 cells x 128 j] h block of dW^T (own thread, so it never
        // delays the ring above) =====================
        if (lane == 0) {
            uint32_t tl = 0;
            TileInfo ti;
            for (int sk = 0; sk < n_seq; ++sk) {
            const int tile = tile_at(sk);
                if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
                const int row0 = (ti.b * tpu + ti.first_cell / kTileM) * kTileM;
                mbar_wait(smem_u32(&bars->hd_empty), (tl & 1) ^ 1);
                const uint32_t full = smem_u32(&bars->hd_full);
                mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), 2 * kABlockBytes);
                for (int t = 0; t < 2; ++t)
                    tma_load_2d_pair(smem_u32(sHd + (size_t)t * kABlockBytes), &tmap_hd, j0_r + t * kBlockK, row0, full);
                ++tl;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            uint32_t slot = 0, sphase = 0, zc = 0, dc = 0;
            const uint32_t w_lo0 = desc_lo(smem_u32(sW), 16);
            const uint32_t hz_lo0 = desc_lo(smem_u32(sHz), 16);
            const uint32_t hd_mn_lo0 = desc_lo(smem_u32(sHd), kABlockBytes);
            const uint32_t dz_lo0 = desc_lo(smem_u32(sDz), 16);
            const uint32_t idesc_z = umma_idesc_bf16(2 * kTileM, kTileM);
            const uint32_t idesc_dw = umma_idesc_bf16(2 * kTileM, 2 * kTileM, 1, 0);
            auto z_mma = [&]() {
                const uint32_t zb = zc & 1;
                mbar_wait(smem_u32(&bars->z_empty[zb]), ((zc >> 1) & 1) ^ 1);
                const uint32_t d_tmem = tmem_base + zb * kBwdChunk;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(smem_u32(&bars->hz_full[slot]), sphase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a_lo = w_lo0 + kb * (kABlockBytes >> 4);
                        const uint32_t b_lo = hz_lo0 + slot * (kHzBytes >> 4);
#pragma unroll
                        for (int k16 = 0; k16 < kBlockK / 16; ++k16)
                            umma_bf16_pair(d_tmem, mk_desc(a_lo + 2 * k16), mk_desc(b_lo + 2 * k16), idesc_z,
                                           (kb | k16) != 0);
                        umma_commit_pair(smem_u32(&bars->hz_empty[slot]));
                        if (kb == KB - 1) umma_commit_pair(smem_u32(&bars->z_full[zb]));
                    }
                    __syncwarp();
                    if (++slot == kHzSlots) { slot = 0; sphase ^= 1; }
                }
                ++zc;
            };
            auto dw_mma = [&]() {
                mbar_wait(smem_u32(&bars->dz_full), dc & 1);
                mbar_wait(smem_u32(&bars->hd_full), dc & 1);
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int kk = 0; kk < kTileM / 16; ++kk)
                        umma_bf16_pair(tmem_acc2, mk_desc(hd_mn_lo0 + kk * (2048 >> 4)),
                                       mk_desc(dz_lo0 + (kk >> 2) * (kABlockBytes >> 4) + (kk & 3) * 2), idesc_dw,
                                       (dc || kk) ? 1u : 0u);
                    umma_commit_pair(smem_u32(&bars->dz_empty));
                    umma_commit_pair(smem_u32(&bars->hd_empty));
                }
                __syncwarp();
                ++dc;
            };
            mbar_wait(smem_u32(&bars->w_full), 0);
            bool have_cur = false;
            TileInfo ti;
            for (int sk = 0; sk < n_seq; ++sk) {
            const int tile = tile_at(sk);
                if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
                z_mma();                    // z^T of this tile runs while the epilogue works on the previous one
                if (have_cur) dw_mma();     // dW^T of the previous tile
                have_cur = true;
            }
            if (have_cur) {
                dw_mma();
                if (elect_one_sync()) umma_commit_pair(smem_u32(&bars->acc_full));
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps): thread = vocab row, half of the cells =====================
        const int e = threadIdx.x - 128;
        const int q = warp & 3, hf = (warp - 4) >> 2;
        const int row = q * 32 + lane;                    // vocab row inside this CTA (TMEM lane)
        const int v = v0_r + row;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float bias2 = (v < V ? __ldg(b_out + v) : 0.f) * kLog2e;
        const bool is_blank = v == blank;
        const uint32_t z_empty_addr[2] = {mapa_shared(smem_u32(&bars->z_empty[0]), 0),
                                          mapa_shared(smem_u32(&bars->z_empty[1]), 0)};
        const uint32_t dz_full_addr = mapa_shared(smem_u32(&bars->dz_full), 0);
        uint8_t* rowp = sDz + hf * kABlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
        uint32_t zc = 0;
        float rowsum = 0.f;
        bool any_tile = false;
        TileInfo ti;
        // per-cell scalars are fetched ONE TILE AHEAD into registers of the 128 staging threads: their
        // dependent global loads (labels, lse, gamma) would otherwise sit on the z -> dz -> dW chain
        int nk = -1;
        RowCtx rc_next;
        auto fetch_next = [&]() {
            TileInfo ni;
            do {
                if (++nk >= n_seq) { nk = n_seq; return; }
            } while (!tile_info<1>(tile_at(nk), tiles_per_utt, 0, tlen, ulen, T, U1, ni));
            if (e < kTileM) load_row_ctx(rc_next, ni, e, T, U1, V, labels, lse, gamma2, grad_cost);
        };
        fetch_next();
        for (int sk = 0; sk < n_seq; ++sk) {
            const int tile = tile_at(sk);
            if (!tile_info<1>(tile, tiles_per_utt, 0, tlen, ulen, T, U1, ti)) continue;
            // ---- per-cell scalars of the tile (loaded during the previous tile)
            if (e < kTileM) {
                s_c2[e] = rc_next.c2; s_gg[e] = rc_next.gg; s_cb[e] = rc_next.corr_b; s_cl[e] = rc_next.corr_l;
                s_lab[e] = rc_next.lab;
            }
            named_bar_sync(1, kEpiThreads);
            fetch_next();
            const uint32_t zb = zc & 1;
            mbar_wait(smem_u32(&bars->z_full[zb]), (zc >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + lane_base + zb * kBwdChunk + hf * 64;
            tmem_ld_32x32b_x32(taddr, r0);
            tmem_ld_32x32b_x32(taddr + 32, r1);
            tmem_wait_ld();
            bool dz_free = false;
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                const uint32_t(&r)[32] = gi == 0 ? r0 : r1;
                const int c0 = hf * 64 + gi * 32;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 c2 = *reinterpret_cast<const float4*>(s_c2 + c0 + i);
                    const float4 gg = *reinterpret_cast<const float4*>(s_gg + c0 + i);
                    float d0 = gg.x * ex2_approx(fmaf(__uint_as_float(r[i + 0]), kLog2e, bias2 + c2.x));
                    float d1 = gg.y * ex2_approx(fmaf(__uint_as_float(r[i + 1]), kLog2e, bias2 + c2.y));
                    float d2 = gg.z * ex2_approx(fmaf(__uint_as_float(r[i + 2]), kLog2e, bias2 + c2.z));
                    float d3 = gg.w * ex2_approx(fmaf(__uint_as_float(r[i + 3]), kLog2e, bias2 + c2.w));
                    if (is_blank) {
                        const float4 cb = *reinterpret_cast<const float4*>(s_cb + c0 + i);
                        d0 -= cb.x; d1 -= cb.y; d2 -= cb.z; d3 -= cb.w;
                    }
                    rowsum += (d0 + d1) + (d2 + d3);
                    pk[i >> 1] = pack_bf16x2(d0, d1);
                    pk[(i >> 1) + 1] = pack_bf16x2(d2, d3);
                }
                if (!dz_free) {  // dW^T of the previous tile must have consumed dz^T
                    mbar_wait(smem_u32(&bars->dz_empty), (zc & 1) ^ 1);
                    dz_free = true;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int chunk = (gi * 4 + j) ^ (row & 7);
                    *reinterpret_cast<uint4*>(rowp + (chunk << 4)) =
                        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            named_bar_sync(2, kEpiThreads);   // dense dz^T complete in this CTA
            // ---- sparse part: cell e subtracts its label term from row (lab - v0_r) if this CTA owns it
            if (e < kTileM) {
                const int lab = s_lab[e];
                const float cl = s_cl[e];
                const int vv = lab - v0_r;
                if (lab >= 0 && cl != 0.f && vv >= 0 && vv < kTileM) {
                    const int c64 = e & 63;
                    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(
                        sDz + (e >> 6) * kABlockBytes + (vv >> 3) * 1024 + (vv & 7) * 128 +
                        (((c64 >> 3) ^ (vv & 7)) << 4)) + (c64 & 7);
                    *p = __float2bfloat16_rn(__bfloat162float(*p) - cl);
                    if (g == 0) atomicAdd(d_b_out + lab, -cl);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(z_empty_addr[zb]);
                mbar_arrive_cluster(dz_full_addr);
            }
            ++zc;
            any_tile = true;
        }
        if (any_tile) {
            // ---- flush dW^T: TMEM lane = j row, columns = the pair's 256 vocab rows
            mbar_wait(smem_u32(&bars->acc_full), 0);
            tc_fence_after();
            const int j = j0_r + row;
            for (int gi = hf * 4; gi < hf * 4 + 4; ++gi) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_acc2 + lane_base + gi * 32, r);
                tmem_wait_ld();
                if (flush_dw) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int vv = v0 + gi * 32 + i;
                        if (vv < V) atomicAdd(d_w_out + (size_t)vv * J + j, __uint_as_float(r[i]));
                    }
                }
            }
            if (g == 0 && v < V) atomicAdd(d_b_out + v, rowsum);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

size_t dwt_smem_bytes(int J) {
    return (size_t)(J / kBlockK) * kABlockBytes + (size_t)kHzSlots * kHzBytes + 2 * kABlockBytes + kDzBytes +
           sizeof(DwtBarriers) + 5 * kTileM * sizeof(float);
}

// Both axis reductions of dpre (B,T,U1,J) bf16 in ONE pass over the tensor:
//   d_enc_proj[b,t,j] = sum_{u <= U_b} dpre[b,t,u,j]   (0 for t >= T_b)        written directly
//   d_dec_proj[b,u,j] = sum_{t <  T_b} dpre[b,t,u,j]   (0 for u >  U_b)        pre-zeroed, red.add.v4
// Block = (128-column slice, kRedTG frames, utterance).  Warp w owns the rows u = w mod 8: for each
// of its u it loads the kRedTG frames at once (8-byte loads, next u prefetched), sums them in
// registers for d_dec (one vector red per (u, lane)) and keeps per-frame partials for d_enc, which
// are combined across the 8 warps through shared memory once at the end.  No barrier in the loop.
constexpr int kRedTG = 8;
constexpr int kRedWarps = 8;
constexpr int kRedCols = 128;   // columns per block: 4 per lane (one 8-byte load of 4 bf16)
__global__ void __launch_bounds__(kRedWarps * 32)
reduce_dpre_kernel(const __nv_bfloat16* __restrict__ dpre, const int* __restrict__ tlen,
                   const int* __restrict__ ulen, int T, int U1, int J, float* __restrict__ d_enc,
                   float* __restrict__ d_dec) {
    __shared__ float4 s_enc[kRedWarps][kRedTG][32];
    const int b = blockIdx.z, j0 = blockIdx.x * kRedCols;
    const int t0 = blockIdx.y * kRedTG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T_b = min(max(tlen[b], 1), T), U1b = min(max(ulen[b], 0), U1 - 1) + 1;
    const int nt = max(0, min(kRedTG, T_b - t0));          // valid frames of this block
    const size_t rs = (size_t)J / 4;                        // row stride in uint2
    const uint2* base = reinterpret_cast<const uint2*>(dpre + (((size_t)b * T + t0) * U1) * J + j0) + lane;
    float e[kRedTG][4];
#pragma unroll
    for (int k = 0; k < kRedTG; ++k) e[k][0] = e[k][1] = e[k][2] = e[k][3] = 0.f;
    auto load_u = [&](int u, uint2 (&v)[kRedTG]) {
#pragma unroll
        for (int k = 0; k < kRedTG; ++k)
            v[k] = (k < nt && u < U1b) ? __ldg(base + ((size_t)k * U1 + u) * rs) : make_uint2(0u, 0u);
    };
    uint2 cur[kRedTG], nxt[kRedTG];
    load_u(warp, cur);
    for (int u = warp; u < U1b; u += kRedWarps) {
        load_u(u + kRedWarps, nxt);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) {
            const float f0 = __uint_as_float(cur[k].x << 16), f1 = __uint_as_float(cur[k].x & 0xffff0000u);
            const float f2 = __uint_as_float(cur[k].y << 16), f3 = __uint_as_float(cur[k].y & 0xffff0000u);
            a0 += f0; a1 += f1; a2 += f2; a3 += f3;
            e[k][0] += f0; e[k][1] += f1; e[k][2] += f2; e[k][3] += f3;
        }
        if (nt > 0) red_add_v4(d_dec + ((size_t)b * U1 + u) * J + j0 + lane * 4, a0, a1, a2, a3);
#pragma unroll
        for (int k = 0; k < kRedTG; ++k) cur[k] = nxt[k];
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

size_t dh_smem_bytes(int J) {
    return (size_t)(J / kBlockK) * kABlockBytes + kDzBytes + (size_t)kDhSlots * kSlotBytes + sizeof(DhBarriers) +
           2 * kBwdChunk * sizeof(float);
}
size_t dw_smem_bytes(int J) {
    return (size_t)(J / kBlockK) * kABlockBytes + kDzBytes + (size_t)kDwSlots * kSlotBytes + sizeof(DwBarriers) +
           kBwdChunk * sizeof(float);
}

}  // namespace

size_t joint_bf16_workspace(int op, int B, int T, int U1, int J, int V) {
    if (op == EMO_OP_RNNT_JOINT_HCACHE) return align_up(hcache_bytes_for(B, T, U1, J), 256);
    if (op == EMO_OP_RNNT_JOINT_HZCACHE) {
        if (!joint_zc_supported(J) || J % 128 != 0 || V % 32 != 0) return 0;
        return zcache_offset_for(B, T, U1, J) + align_up(zcache_bytes_for(B, T, U1, V), 256);
    }
    size_t w = align_up((size_t)V * J * sizeof(__nv_bfloat16), 256);
    if (op == EMO_OP_RNNT_JOINT_BWD)   // dpre (B,T,U1,J) (recompute path) or tile-major dh (z-cache path, >= as large)
        return w + align_up(hcache_bytes_for(B, T, U1, J), 256);
    // forward: + fp16 copies of enc_proj and dec_proj
    return w + align_up((size_t)B * T * J * sizeof(__half), 256) + align_up((size_t)B * U1 * J * sizeof(__half), 256);
}

int joint_bf16_launches(int op, int B, int T, int U1, int J, int V) {
    (void)B; (void)T; (void)U1; (void)J; (void)V;
    if (op == EMO_OP_RNNT_JOINT_BWD) return 4;  // weight cast, dh kernel, axis reductions, dW kernel
    return 4;                                    // weight cast, 2 stream casts, fused joint forward
}

int joint_bwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen,
                   const float* lse, const float* gamma2, const float* grad_cost,
                   const void* hcache, size_t hcache_bytes, int B, int T,
                   int U1, int J, int V, int blank, float* d_enc_proj, float* d_dec_proj,
                   float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    EMO_REQUIRE(w_out && b_out && labels && tlen && ulen && lse && gamma2 && grad_cost && d_enc_proj &&
                    d_dec_proj && d_w_out && d_b_out && ws,
                EMO_BAD_ARG, "joint_bwd(bf16): null pointer");
    EMO_REQUIRE(hcache, EMO_BAD_ARG, "joint_bwd(bf16): the h cache written by emo_rnnt_joint_fwd is required");
    int rc = check_bf16_shape(B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(hcache_bytes >= hcache_bytes_for(B, T, U1, J) && ((uintptr_t)hcache & 255) == 0,
                EMO_WORKSPACE_TOO_SMALL, "joint_bwd(bf16): h cache too small or misaligned");
    EMO_REQUIRE(ws_bytes >= joint_bf16_workspace(EMO_OP_RNNT_JOINT_BWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_bwd(bf16): workspace too small");
    EMO_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)w_out & 15) == 0 && ((uintptr_t)d_w_out & 15) == 0,
                EMO_BAD_ARG, "joint_bwd(bf16): pointers must be 16-byte (workspace 256-byte) aligned");
    __nv_bfloat16* w_bf16 = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* dpre = reinterpret_cast<__nv_bfloat16*>(
        (char*)ws + align_up((size_t)V * J * sizeof(__nv_bfloat16), 256));
    const size_t nw = (size_t)V * J;
    f32_to_bf16_kernel<<<ceil_div(nw, 4 * 256), 256, 0, st>>>(w_out, w_bf16, nw);
    EMO_CHECK_LAUNCH("f32_to_bf16_kernel");
    EMO_CUDA(cudaMemsetAsync(d_w_out, 0, nw * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_out, 0, (size_t)V * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_dec_proj, 0, (size_t)B * U1 * J * sizeof(float), st));

    // ---- z-cache path: the forward left the logits behind the h cache
    if (joint_zc_supported(J) &&
        hcache_bytes >= zcache_offset_for(B, T, U1, J) + zcache_bytes_for(B, T, U1, V)) {
        const void* zcache = (const char*)hcache + zcache_offset_for(B, T, U1, J);
        rc = joint_dhz_launch(w_bf16, hcache, zcache, labels, tlen, ulen, lse, gamma2, grad_cost, B, T, U1, J, V,
                              blank, dpre, enc_proj, dec_proj, d_enc_proj, d_dec_proj, st);
        if (rc) return rc;
        return joint_dwz_launch(hcache, zcache, labels, tlen, ulen, lse, gamma2, grad_cost, B, T, U1, J, V, blank,
                                d_w_out, d_b_out, st);
    }

    // ---- recompute path (h cache only)
    const int KB = J / kBlockK;
    const int NCH = ceil_div(V, kBwdChunk), NPART = ceil_div(KB, kPartBlocks);
    EMO_REQUIRE(NCH * NPART <= sm_count() * 4, EMO_UNSUPPORTED_SHAPE,
                "joint_bwd(bf16): vocabulary %d too large for the dW role grid", V);
    CUtensorMap tmap_wz, tmap_wd, tmap_h;
    rc = make_tmap_bf16_2d(&tmap_wz, w_bf16, (uint64_t)J, (uint64_t)V, kBlockK, 64);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmap_wd, w_bf16, (uint64_t)J, (uint64_t)V, kBlockK, 128);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmap_h, hcache, (uint64_t)J, (uint64_t)B * tiles128_per_utt(T, U1) * kTileM, kBlockK,
                           kTileM);
    if (rc) return rc;

    // ---- dh kernel (CTA pairs) + the two axis reductions
    {
        const size_t smem = dh_smem_bytes(J);
        EMO_REQUIRE(smem <= (size_t)kSmemLimit, EMO_UNSUPPORTED_SHAPE, "joint_bwd(bf16): shared memory (dh)");
        EMO_CUDA(cudaFuncSetAttribute(joint_dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int ptiles = B * ceil_div((size_t)T * U1, 2 * kTileM);
        const int pairs = max(1, min(ptiles, sm_count() / 2));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        EMO_CUDA(cudaLaunchKernelEx(&cfg, joint_dh_kernel, tmap_wz, tmap_wd, tmap_h, b_out, labels, tlen, ulen, lse,
                                    gamma2, grad_cost, B, T, U1, J, V, blank, dpre));
        EMO_CHECK_LAUNCH("joint_dh_kernel");
    }
    {
        reduce_dpre_kernel<<<dim3(J / kRedCols, ceil_div(T, kRedTG), B), kRedWarps * 32, 0, st>>>(
            dpre, tlen, ulen, T, U1, J, d_enc_proj, d_dec_proj);
        EMO_CHECK_LAUNCH("reduce_dpre_kernel");
    }

    // ---- dW kernel
    static const bool use_dw_v1 = getenv("EMO_DW_V1") != nullptr;  // A/B switch: per-CTA roles, streamed W
    if (!use_dw_v1 && dwt_smem_bytes(J) <= (size_t)kSmemLimit) {
        // transposed CTA-pair roles (256 vocab rows x 256 hidden units), W resident
        CUtensorMap tmap_hz;
        rc = make_tmap_bf16_2d(&tmap_hz, hcache, (uint64_t)J, (uint64_t)B * tiles128_per_utt(T, U1) * kTileM,
                               kBlockK, 64);
        if (rc) return rc;
        const size_t smem = dwt_smem_bytes(J);
        EMO_CUDA(cudaFuncSetAttribute(joint_dwt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = B * ceil_div((size_t)T * U1, kTileM);
        const int roles = ceil_div(J, 256) * ceil_div(V, 256);
        const int splits = max(1, min((sm_count() / 2) / roles, tiles));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * roles * splits);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        EMO_CUDA(cudaLaunchKernelEx(&cfg, joint_dwt_kernel, tmap_wd, tmap_hz, tmap_h, b_out, labels, tlen, ulen, lse,
                                    gamma2, grad_cost, B, T, U1, J, V, blank, splits, d_w_out, d_b_out));
        EMO_CHECK_LAUNCH("joint_dwt_kernel");
    } else {
        const size_t smem = dw_smem_bytes(J);
        EMO_REQUIRE(smem <= (size_t)kSmemLimit, EMO_UNSUPPORTED_SHAPE, "joint_bwd(bf16): shared memory (dW)");
        EMO_CUDA(cudaFuncSetAttribute(joint_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = B * ceil_div((size_t)T * U1, kTileM);
        const int roles = NCH * NPART;
        const int splits = max(1, min(sm_count() / roles, tiles));
        joint_dw_kernel<<<roles * splits, kThreads, smem, st>>>(tmap_wd, tmap_h, b_out, labels, tlen, ulen, lse,
                                                                gamma2, grad_cost, B, T, U1, J, V, blank, splits,
                                                                d_w_out, d_b_out);
        EMO_CHECK_LAUNCH("joint_dw_kernel");
    }
    return EMO_OK;
}

}  // namespace emo

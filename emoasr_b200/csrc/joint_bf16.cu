// EMO_PREC_BF16: fused joint on the 5th-gen tensor cores.
//
//   z[cell, v] = sum_j tanh(enc[b,t,j] + dec[b,u,j]) * w_out[v,j] + b_out[v]
//   lse[cell]  = log sum_v exp z ;  lp2[cell] = { z[blank]-lse, z[label]-lse }
// (asr/modeling/decoders/rnn_transducer.py:147-156 + :102).  z lives only in TMEM and registers: nothing of size
// N x V is written to memory.
//
// Persistent, warp-specialised CTA pairs (cluster of 2, cta_group::2), one 256-cell tile per pair at a time (cells of
// one utterance, flattened over its VALID (t,u) region, so padding costs nothing); 640 threads, setmaxnreg budgets
// per warpgroup:
//   warp 0       TMA producer   w_out (bf16, [V][J]) tiles [128 v x 64 j] per CTA, 128B swizzle, 5-stage mbarrier
//                               ring, complete_tx
//   warp 1       MMA issuer     tcgen05.mma kind::f16, M=256, N<=256, K=16; accumulators in TMEM, two 256-column
//                               buffers so the epilogue of vocab chunk n overlaps the MMAs of chunk n+1
//   warp 2       TMEM allocator
//   warps 4-11   epilogue       two warps per TMEM lane quadrant, each half of a chunk's columns: tcgen05.ld 32
//                               columns at a time (thread == lattice cell), bias add, online (max, sum-exp) over the
//                               vocab chunks, capture of the blank and label logits
//   warps 12-19  A producers    h = tanh(enc+dec): fp16 gathers, packed-half add and tanh.approx, -> bf16 -> shared
//                               memory in the canonical K-major 128B-swizzle layout, one 64-wide K block at a time so
//                               the MMAs of the next tile start as soon as block 0 is rewritten (8 warps, 16 rows
//                               each: with 4 the MMA issuer spent a third of its time waiting for h)
// The h tile (128 x J bf16 per CTA) stays resident in shared memory for all vocab chunks of the tile.
// (Round 1 also had a variant that stored z as fp16 and h as bf16 for a "z-cache" backward.  It materialised the
// logits the design exists to avoid and its backward corrupted one tile in ~2.5 % of the runs at the cfg-4 shape
// (tools/stress_repeat.py); removed in round 2 in favour of the ring backward, joint_bwd_ring.cu.)
#include "joint_tc.cuh"

// -DEMO_ZC_PROF: clock64 accounting of the MMA issuer's mbarrier waits (printf from CTA 0); tools/gpu_zcprof.sh
#ifdef EMO_ZC_PROF
#define EMO_PROF(...) __VA_ARGS__
#else
#define EMO_PROF(...)
#endif

namespace emo {
namespace {

constexpr int kCtas = 2;            // CTAs per tile (cluster of 2, cta_group::2)
constexpr int kBStages = 5;         // w_out ring
constexpr int kBRows = kChunkN / 2; // vocab rows of a w_out tile held by one CTA

struct __align__(16) FwdBarriers {
    uint64_t b_full[kBStages], b_empty[kBStages];
    uint64_t a_full[kMaxKBlocks], a_empty[kMaxKBlocks];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

// One 32-column group of logits of this thread's row: bias add, online (max, sum exp2), capture of
// the blank / label logit.
__device__ __forceinline__ void lse_group(const uint32_t (&r)[32], const float* __restrict__ bias,
                                          int v0, int lab, int blank, float& run_m, float& run_s,
                                          float& zb, float& zl) {
    // packed fp32 pairs (FADD2 / FFMA2) for the bias add, the exponent argument and the sums: 3.5 instructions per
    // logit with the max and the MUFU instead of 5 -- the epilogue warps, not the MMAs, set the pace of this kernel
    float x[32];
    float m0 = kNegInf, m1 = kNegInf, m2 = kNegInf, m3 = kNegInf;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(bias + i);
        const float2 a = __fadd2_rn(make_float2(__uint_as_float(r[i + 0]), __uint_as_float(r[i + 1])), make_float2(bv.x, bv.y));
        const float2 b = __fadd2_rn(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), make_float2(bv.z, bv.w));
        x[i + 0] = a.x; x[i + 1] = a.y; x[i + 2] = b.x; x[i + 3] = b.y;
        m0 = fmaxf(m0, a.x);
        m1 = fmaxf(m1, a.y);
        m2 = fmaxf(m2, b.x);
        m3 = fmaxf(m3, b.y);
    }
    const float new_m = fmaxf(run_m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    const float neg_m2 = -new_m * kLog2e;
    const float2 nm = make_float2(neg_m2, neg_m2), l2 = make_float2(kLog2e, kLog2e);
    float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float2 t0 = __ffma2_rn(make_float2(x[i + 0], x[i + 1]), l2, nm);
        const float2 t1 = __ffma2_rn(make_float2(x[i + 2], x[i + 3]), l2, nm);
        s01 = __fadd2_rn(s01, make_float2(ex2_approx(t0.x), ex2_approx(t0.y)));
        s23 = __fadd2_rn(s23, make_float2(ex2_approx(t1.x), ex2_approx(t1.y)));
    }
    run_s = run_s * ex2_approx((run_m - new_m) * kLog2e) + ((s01.x + s01.y) + (s23.x + s23.y));
    run_m = new_m;
    const int dl = lab - v0;
    const bool mine = dl >= 0 && dl < 32;
    if (__any_sync(0xffffffffu, mine)) {
        const float v = mux32(x, dl & 31);
        zl = mine ? v : zl;
    }
    if (blank >= v0 && blank < v0 + 32) zb = mux32(x, blank - v0);  // warp-uniform branch
}

// =================================================================================================
// Fused joint forward.  See the file header for the role layout.  The leader CTA of the pair issues M=256 MMAs over
// both CTAs' h tiles, and each CTA stages only half (128 vocab rows) of every w_out tile.  Barrier topology:
//   b_full / a_full / acc_empty live in the LEADER (arrivals from both CTAs, TMA bytes from both),
//   b_empty / a_empty / acc_full are signalled in BOTH CTAs by a multicast tcgen05.commit.
constexpr int kFwdThreads = 640;  // 4 control warps, 8 epilogue warps, 8 A-producer warps
constexpr int kFwdEpiThreads = 256;
constexpr int kFwdProdWarps = 8;
constexpr int kBBytes = kBRows * kBlockK * 2;
template <int N> __device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// A-operand producer: this warp's 16 rows of the 128-row tile, one 64-wide K block.
// enc / dec arrive as fp16 (one 16-byte load = 8 hidden units per row and stream), h = tanh(enc + dec)
// with packed-half add and tanh.approx.f16x2, rounded to bf16 into the canonical K-major SW128 layout
// (16-byte chunk index XOR (row mod 8)).
__device__ __forceinline__ void produce_h_block16(const uint4 (&re)[4], const uint4 (&rd)[4], int pw, int rsub,
                                                  int c, uint8_t* blk, bool plain) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int row = pw * 16 + p * 4 + rsub;
        const uint32_t e[4] = {re[p].x, re[p].y, re[p].z, re[p].w};
        const uint32_t d[4] = {rd[p].x, rd[p].y, rd[p].z, rd[p].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_f16x2(plain ? e[q] : tanh_f16x2(hadd2_u32(e[q], d[q])));
            o[q] = pack_bf16x2(f.x, f.y);
        }
        uint8_t* dst = blk + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// kPlain (the CTC head, ctc_head.cu): the A operand is the enc stream itself (h = enc, no dec stream, no tanh);
// the caller passes U1 = 1 and ulen = 0, so a "cell" is a frame.
template <bool kPlain>
__global__ void __launch_bounds__(kFwdThreads, 1)
joint_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __half* __restrict__ enc,
                 const __half* __restrict__ dec, const float* __restrict__ b_out,
                 const int* __restrict__ labels, const int* __restrict__ tlen,
                 const int* __restrict__ ulen, int B, int T, int U1, int J, int V, int blank,
                 float* __restrict__ lp2, float* __restrict__ lse_out) {
    using Bars = FwdBarriers;
    // 1024-byte alignment is required by the 128B swizzle atoms; the kernel has no static shared
    // memory, so the dynamic window starts at the CTA's (1 KiB-granular) shared-memory base.
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = J / kBlockK;
    const int NC = (V + kChunkN - 1) / kChunkN;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)KB * kABlockBytes;
    Bars* bars = reinterpret_cast<Bars*>(sB + (size_t)kBStages * kBBytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);  // [2][kChunkN]
    float4* s_part = reinterpret_cast<float4*>(s_bias + 2 * kChunkN);  // [kTileM] partial LSE state of column half 1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_per_utt = (T * U1 + kCtas * kTileM - 1) / (kCtas * kTileM);
    const int total_tiles = B * tiles_per_utt;
    const int tile0 = blockIdx.x / kCtas, tile_stride = gridDim.x / kCtas;
    constexpr uint32_t kArrivals = (kFwdEpiThreads / 32) * kCtas;   // epilogue WARPS of the pair (one arrive each)

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(smem_u32(&bars->b_full[i]), kCtas);
            mbar_init(smem_u32(&bars->b_empty[i]), 1);
        }
        for (int i = 0; i < kMaxKBlocks; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), kFwdProdWarps * kCtas);   // producer WARPS of the pair
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), kArrivals);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap_w);
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // register budget per warpgroup (x128 threads): control 40, two epilogue groups 128, two producer groups 88
    // (two units of gather loads in flight) = 472 of the 480 the CTA owns at launch (96 x 640); an exact fit
    // deadlocks in setmaxnreg.inc
    if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            TileInfo ti;
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<kCtas>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
                for (int nc = 0; nc < NC; ++nc) {
                    const int n = min(kChunkN, V - nc * kChunkN);
                    const int y = nc * kChunkN + (int)rank * (n >> 1);
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(smem_u32(&bars->b_empty[stage]), phase ^ 1);
                        const uint32_t full = smem_u32(&bars->b_full[stage]);
                        mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), kBBytes);
                        tma_load_2d_pair(smem_u32(sB + (size_t)stage * kBBytes), &tmap_w, kb * kBlockK, y, full);
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        if (leader) {
            uint32_t stage = 0, phase = 0, cc = 0, tl = 0;
            constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
            TileInfo ti;
            EMO_PROF(long long p_acc = 0, p_a = 0, p_b = 0, p_t0 = clock64(), p_c;)
            for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
                if (!tile_info<kCtas>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
                for (int nc = 0; nc < NC; ++nc, ++cc) {
                    const uint32_t buf = cc & 1;
                    EMO_PROF(p_c = clock64();)
                    mbar_wait_cluster(smem_u32(&bars->acc_empty[buf]), ((cc >> 1) & 1) ^ 1);
                    EMO_PROF(p_acc += clock64() - p_c;)
                    const int n = min(kChunkN, V - nc * kChunkN);
                    const uint32_t idesc = umma_idesc_bf16(kCtas * kTileM, n);
                    const uint32_t d_tmem = tmem_base + buf * kChunkN;
                    for (int kb = 0; kb < KB; ++kb) {
                        EMO_PROF(p_c = clock64();)
                        if (nc == 0) mbar_wait_cluster(smem_u32(&bars->a_full[kb]), tl & 1);
                        EMO_PROF(p_a += clock64() - p_c; p_c = clock64();)
                        mbar_wait_cluster(smem_u32(&bars->b_full[stage]), phase);
                        EMO_PROF(p_b += clock64() - p_c;)
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t a_lo = a_lo0 + kb * (kABlockBytes >> 4);
                            const uint32_t b_lo = b_lo0 + stage * (kBBytes >> 4);
#pragma unroll
                            for (int k16 = 0; k16 < kBlockK / 16; ++k16) {
                                const uint64_t ad = ((uint64_t)kDescHi << 32) | (a_lo + 2 * k16);
                                const uint64_t bd = ((uint64_t)kDescHi << 32) | (b_lo + 2 * k16);
                                umma_bf16_pair(d_tmem, ad, bd, idesc, (kb | k16) != 0);
                            }
                            umma_commit_pair(smem_u32(&bars->b_empty[stage]));
                            if (nc == NC - 1) umma_commit_pair(smem_u32(&bars->a_empty[kb]));
                            if (kb == KB - 1) umma_commit_pair(smem_u32(&bars->acc_full[buf]));
                        }
                        __syncwarp();
                        if (++stage == kBStages) { stage = 0; phase ^= 1; }
                    }
                }
                ++tl;
            }
            EMO_PROF(if (blockIdx.x == 0 && lane == 0)
                         printf("fwd issuer: total %lld clk, %u tiles; wait acc_empty %lld a_full %lld b_full %lld\n",
                                clock64() - p_t0, tl, p_acc, p_a, p_b);)
        }
    }
    } else if (warp < 12) {
        reg_inc<128>();
        // ===================== epilogue: online LSE over vocab chunks =====================
        // Two warps per TMEM lane quadrant: warp (q, hf) owns columns [128 hf, 128 hf + 128) of every 256-wide
        // chunk for the rows of quadrant q.  The two partial (max, sum, blank, label) states of a row are
        // merged through shared memory once per tile.
        const int q = warp & 3, hf = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int etid = threadIdx.x - 128;   // 0..255
        uint32_t cc = 0;
        TileInfo ti;
        const uint32_t acc_empty0 = mapa_shared(smem_u32(&bars->acc_empty[0]), 0);
        float nb = etid < V ? __ldg(b_out + etid) : 0.f;   // bias of the first chunk of the first tile
        for (int tile = tile0; tile < total_tiles; tile += tile_stride) {
            if (!tile_info<kCtas>(tile, tiles_per_utt, rank, tlen, ulen, T, U1, ti)) continue;
            const int m = ti.first_cell + row;
            const bool valid = m < ti.n_cells;
            const int t = valid ? m / ti.U1b : 0;
            const int u = valid ? m - t * ti.U1b : 0;
            int lab = -1;
            if (valid && u < ti.U1b - 1)
                lab = min(max(__ldg(labels + (size_t)ti.b * (U1 - 1) + u), 0), V - 1);
            float run_m = kNegInf, run_s = 0.f, zb = 0.f, zl = 0.f;
            for (int nc = 0; nc < NC; ++nc, ++cc) {
                const uint32_t buf = cc & 1;
                const int n = min(kChunkN, V - nc * kChunkN);
                float* bias = s_bias + buf * kChunkN;
                bias[etid] = nb;
                {   // prefetch the next chunk's bias (wraps to chunk 0 for the next tile)
                    const int nn = (nc + 1 == NC) ? 0 : nc + 1;
                    const int i0 = nn * kChunkN + etid;
                    nb = i0 < V ? __ldg(b_out + i0) : 0.f;
                }
                named_bar_sync(1, kFwdEpiThreads);
                mbar_wait(smem_u32(&bars->acc_full[buf]), (cc >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kChunkN;
                const int g0 = hf * 4, g1 = min(hf * 4 + 4, n >> 5);   // my 32-column groups of the chunk
                if (g0 < g1) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(taddr + g0 * 32, ra);
                    for (int g = g0; g < g1; g += 2) {
                        tmem_wait_ld();
                        if (g + 1 < g1) tmem_ld_32x32b_x32(taddr + (g + 1) * 32, rb);
                        lse_group(ra, bias + g * 32, nc * kChunkN + g * 32, lab, blank, run_m, run_s, zb, zl);
                        if (g + 1 < g1) {
                            tmem_wait_ld();
                            if (g + 2 < g1) tmem_ld_32x32b_x32(taddr + (g + 2) * 32, ra);
                            lse_group(rb, bias + (g + 1) * 32, nc * kChunkN + (g + 1) * 32, lab, blank, run_m, run_s,
                                      zb, zl);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty0 + buf * 8);
            }
            // ---- merge the two column halves of the row (half 1 -> shared memory -> half 0)
            if (hf == 1) s_part[row] = make_float4(run_m, run_s, zb, zl);
            named_bar_sync(2, kFwdEpiThreads);
            if (hf == 0 && valid) {
                const float4 o = s_part[row];
                lse_merge_exp2(run_m, run_s, o.x, o.y);
                if (((blank % kChunkN) >> 7) == 1) zb = o.z;
                if (lab >= 0 && ((lab % kChunkN) >> 7) == 1) zl = o.w;
                const float l = run_m + kLn2 * log2f(run_s);
                const size_t cell = ((size_t)ti.b * T + t) * U1 + u;
                reinterpret_cast<float2*>(lp2)[cell] = make_float2(zb - l, lab >= 0 ? zl - l : 0.f);
                lse_out[cell] = l;
            }
        }
    } else {
        reg_dec<88>();
        // ===================== A producers =====================
        const int pw = warp - 12;      // 8 producer warps, 16 rows each
        const int c = lane & 7;        // 16-byte chunk (8 hidden units) inside the 64-wide K block
        const int rsub = lane >> 3;    // 4 rows per warp pass
        // Flattened (tile, K block) sequence with TWO units of loads in flight: the loads of unit n+2 are issued
        // right after unit n has been written, so an L2 round trip is hidden behind two unit periods (and, across
        // tiles, behind the wait for the MMAs to release the slot).
        int ltile = tile0 - tile_stride, lunit = KB;   // load cursor; unit = kb
        uint32_t eoff[4], doff[4];
        auto issue = [&](uint4 (&re)[4], uint4 (&rd)[4]) -> bool {
            if (lunit == KB) {
                TileInfo ti;
                do {
                    ltile += tile_stride;
                    if (ltile >= total_tiles) { ltile = total_tiles; return false; }
                } while (!tile_info<kCtas>(ltile, tiles_per_utt, rank, tlen, ulen, T, U1, ti));
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int row = pw * 16 + p * 4 + rsub;
                    const int m = min(ti.first_cell + row, ti.n_cells - 1);  // clamp padding rows
                    const int t = m / ti.U1b, u = m - t * ti.U1b;
                    eoff[p] = (uint32_t)(((size_t)ti.b * T + t) * J) + c * 8;
                    doff[p] = (uint32_t)(((size_t)ti.b * U1 + u) * J) + c * 8;
                }
                lunit = 0;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                re[p] = __ldg(reinterpret_cast<const uint4*>(enc + eoff[p] + lunit * kBlockK));
                rd[p] = kPlain ? make_uint4(0u, 0u, 0u, 0u)
                               : __ldg(reinterpret_cast<const uint4*>(dec + doff[p] + lunit * kBlockK));
            }
            ++lunit;
            return true;
        };
        uint32_t wkb = 0, wtl = 0;                 // work cursor: same sequence, KB units per valid tile
        const uint32_t a_full0 = mapa_shared(smem_u32(&bars->a_full[0]), 0);
        EMO_PROF(long long q_wait = 0, q_work = 0, q_t0 = clock64(), q_c;)
        auto work = [&](const uint4 (&re)[4], const uint4 (&rd)[4]) {
            EMO_PROF(q_c = clock64();)
            mbar_wait(smem_u32(&bars->a_empty[wkb]), (wtl & 1) ^ 1);
            EMO_PROF(q_wait += clock64() - q_c; q_c = clock64();)
            produce_h_block16(re, rd, pw, rsub, c, sA + (size_t)wkb * kABlockBytes, kPlain);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_full0 + wkb * 8);
            EMO_PROF(q_work += clock64() - q_c;)
            if (++wkb == (uint32_t)KB) { wkb = 0; ++wtl; }
        };
        uint4 e0[4], d0[4], e1[4], d1[4];
        bool v0 = issue(e0, d0);
        bool v1 = v0 && issue(e1, d1);
        while (v0) {
            work(e0, d0);
            v0 = v1 && issue(e0, d0);
            if (!v1) break;
            work(e1, d1);
            v1 = v0 && issue(e1, d1);
        }
        EMO_PROF(if (blockIdx.x == 0 && threadIdx.x == 12 * 32)
                     printf("fwd producer warp 12: total %lld clk, %u tiles; wait a_empty %lld, produce (incl. load stalls) "
                            "%lld, rest (issue) %lld\n", clock64() - q_t0, wtl, q_wait, q_work,
                            clock64() - q_t0 - q_wait - q_work);)
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace

// the kernel on prepared operands: bf16 w_out (Vp rows), fp16 streams, bias of Vp entries (joint_bf16_casts)
int joint_fwd_launch(const void* w_bf16, const void* enc_h, const void* dec_h, const float* b_pad, const int* labels,
                     const int* tlen, const int* ulen, int B, int T, int U1, int J, int Vp, int blank, float* lp2,
                     float* lse, int plain, cudaStream_t st) {
    const int KB = J / kBlockK;
    CUtensorMap tmap;
    int rc = make_tmap_bf16_2d(&tmap, w_bf16, (uint64_t)J, (uint64_t)Vp, kBlockK, kBRows);
    if (rc) return rc;
    const int tiles = B * ceil_div((size_t)T * U1, kCtas * kTileM);
    const int ctas = kCtas * min(tiles, sm_count() / kCtas);
    const size_t smem = (size_t)KB * kABlockBytes + (size_t)kBStages * kBBytes + sizeof(FwdBarriers) +
                        2 * kChunkN * sizeof(float) + kTileM * sizeof(float4);
    EMO_REQUIRE(smem <= (size_t)kSmemLimit, EMO_UNSUPPORTED_SHAPE, "joint_fwd(bf16): shared memory");
    auto kern = plain ? joint_fwd_kernel<true> : joint_fwd_kernel<false>;
    EMO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kFwdThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EMO_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, (const __half*)enc_h, (const __half*)dec_h, b_pad, labels,
                                tlen, ulen, B, T, U1, J, Vp, blank, lp2, lse));
    EMO_CHECK_LAUNCH("joint_fwd_kernel");
    return EMO_OK;
}

int joint_fwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen, int B,
                   int T, int U1, int J, int V, int blank, float* lp2, float* lse, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
    EMO_REQUIRE(enc_proj && dec_proj && w_out && b_out && labels && tlen && ulen && lp2 && lse && ws,
                EMO_BAD_ARG, "joint_fwd(bf16): null pointer");
    int rc = check_bf16_shape(B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(ws_bytes >= joint_bf16_workspace(EMO_OP_RNNT_JOINT_FWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_fwd(bf16): workspace too small");
    EMO_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)enc_proj & 15) == 0 &&
                    ((uintptr_t)dec_proj & 15) == 0 && ((uintptr_t)w_out & 15) == 0,
                EMO_BAD_ARG, "joint_fwd(bf16): pointers must be 16-byte (workspace 256-byte) aligned");
    const void *w_bf16, *enc_h, *dec_h;
    const float* b_pad;
    rc = joint_bf16_casts(enc_proj, dec_proj, w_out, b_out, B, T, U1, J, V, ws, &w_bf16, &enc_h, &dec_h, &b_pad, st);
    if (rc) return rc;
    // pad columns of the vocabulary: zero weights, bias -1e30 (contribute nothing to the LSE)
    return joint_fwd_launch(w_bf16, enc_h, dec_h, b_pad, labels, tlen, ulen, B, T, U1, J, padded_vocab(V), blank, lp2,
                            lse, 0, st);
}

}  // namespace emo

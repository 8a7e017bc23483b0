// EMO_PREC_BF16 backward of the fused joint WITHOUT any N x V tensor in HBM ("ring" route, the default).
//
//   dz[cell,v] = g_b * (gamma * exp(z - lse) - gamma_blank 1[v=blank] - gamma_label 1[v=label])
//   dW = dz^T h      dh = dz W      (rnn_transducer.py:101-102,147-156 differentiated)
//
// The logits are RECOMPUTED tile by tile on the tensor cores (1 GEMM unit), turned into dz in the epilogue
// and consumed by the two gradient GEMMs (1 unit each): 3 executed units for 2 algorithmic ones, against 6
// for a design that recomputes z separately for every accumulator that does not fit next to it in TMEM
// (a [128 cells x 512] fp32 accumulator IS the 512 TMEM columns of an SM, so neither dh nor dW can share an
// SM with the logit accumulators).  The three GEMMs therefore run on DIFFERENT SMs of one persistent,
// heterogeneous kernel and hand dz / h to each other through a small ring of tiles in global memory that is
// sized to stay L2-resident (36 MB, independent of the problem size):
//
//   P pairs  (producer)   the forward's pipeline again: h = tanh(enc+dec) -> smem, z = h W^T in TMEM
//                         (cta_group::2, 256 cells per pair tile); epilogue thread == lattice cell:
//                         dz -> bf16 -> per-warp staging -> TMA store into the ring; h blocks -> ring.
//   D pairs  (dh)         cell-stationary: dh[256 x J] += dz[256 x 64 v] W[64 v x J] over the vocabulary, dz
//                         straight from the ring by TMA (K-major A operand); drain -> bf16 dh, tile-major
//                         (HBM; the axis reduction kernel applies (1 - h^2) and forms d_enc / d_dec).
//   W pairs  (dW)         vocab-stationary: role = 256 vocab rows, all of J in TMEM, accumulated over the
//                         tiles of the pair's split: dW[256 v x J] += dz^T[256 v x 64 cells] h[64 cells x J];
//                         dz (MN-major A) and h (MN-major B) from the ring; column sums of dz -> d_b_out.
//
// Ring protocol (global int counters, zeroed per launch by ring_prep_kernel): items are produced and consumed
// in increasing order of the dense pair-tile index q (prefix sums over the utterances' valid cells), slot =
// item mod slots.  A writer signals "ready" with red.release.gpu after its TMA stores have completed
// (cp.async.bulk.wait_group 0); a consumer polls (relaxed loads, then one ld.acquire.gpu) before its TMA loads and
// signals "done" once the loads have landed (mbarrier complete_tx observed); a producer waits for "done" of the
// previous use of the slot before overwriting it.  In a producer CTA both global-memory sides of the protocol
// belong to one "ring manager" lane (warp 2): the epilogue warps hand it finished items / receive free slots
// through shared-memory mbarriers.  No role ever waits for a LATER item, all CTAs are co-resident (grid <=
// resident clusters, checked on the host), hence no deadlock.
#include "joint_tc.cuh"

// -DEMO_ZC_PROF: clock64 accounting of what each role's MMA issuer / ring gate waits for (printf from the first
// pair of each role); tools/gpu_ringprof.sh.  Never defined in the shipped library.
#ifdef EMO_ZC_PROF
#define EMO_PROF(...) __VA_ARGS__
#else
#define EMO_PROF(...)
#endif

namespace emo {
namespace {

constexpr int kRingThreads = 640;
constexpr int kVG = 1024;                      // vocab entries per ring item (one z slot = 256 cells x kVG)
constexpr int kPairM = 2 * kTileM;             // cells per pair tile
constexpr int kRingSlots = 48;                 // z slots (512 KiB each at V >= 1024) and h slots (256 KiB at J = 512)
constexpr int kPBStages = 3;                   // P: w_out ring (the dz staging buffers take the fourth stage)
constexpr int kDZStages = 5, kDOpStages = 3;   // D: dz ring, w_out ring
constexpr int kWZStages = 5, kWOpStages = 4;   // W: dz ring, h ring
constexpr int kStageBytes = 16384;
constexpr int kBoxBytes = 8192;                // [64 rows x 128 B]
constexpr int kZStBytes = 2048;                // P epilogue staging block: [32 cells x 32 v] bf16, 64B swizzle
constexpr int kZStBufs = 2;                    // blocks per epilogue warp: a TMA store in flight while the next is built
constexpr int kDrainWarps = 16, kDrainBufBytes = 2048;   // D: warps 4-19, [32 cells x 32 j] bf16 blocks, 64B swizzle
constexpr int kColsumWarps = 8;
constexpr int kMaxB = 1024;                    // utterances (prefix table lives in shared memory)
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO=1024, v1, SW128

struct RingArgs {
    const __half* enc;          // (B,T,J) fp16 copy of enc_proj
    const __half* dec;          // (B,U1,J)
    const float* b_out;
    const int* labels;
    const float* lse;
    const float* lp2;
    const float* gamma2;
    const float* grad_cost;
    const float* grad_lse;      // (B,T,U1) gradient w.r.t. the forward's lse output, or NULL: adds grad_lse * softmax(z) to dz
    const int* prefix;          // [B+1] dense pair-tile prefix, then [B] packed (T_b | U1b << 16)
    int* flags;                 // z_ready[NZ] z_done[NZ] h_ready[NH] h_done[NH]
    float* d_w_out;
    float* d_b_out;
    int B, T, U1, J, V, blank;   // V = vocabulary padded to a multiple of 32 (pad rows of w_out are 0, pad bias -1e30)
    int Vout;                    // rows of d_w_out / entries of d_b_out (the caller's vocabulary)
    int plain;                   // CTC head (ctc_head.cu): h = enc (no dec stream, no tanh), dz = g softmax(z) with a
                                 // PER-CELL grad_cost (B,T,U1), no blank / label patches
    int nP, nD, nS;             // pairs: producers, dh consumers, dW splits (x roles_v pairs)
    int NZ, NH, G;              // ring slots, vocab groups per tile
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* p) {
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// consumer / producer side of a ring flag: wait, then order the async-proxy accesses (TMA) that follow.  The poll
// itself is a relaxed load: an acquire load per iteration invalidates the SM's L1 every time (CCTL.IVALL, 8.4 M of
// them per launch in the first version -- the epilogue's register spills and the producers' gathers then miss L1);
// one acquire load of the value that passed does the synchronisation.
__device__ __forceinline__ bool ring_try(const int* p, int need) {
    if (ld_relaxed_gpu(p) < need) return false;
    (void)ld_acquire_gpu(p);
    fence_proxy_async_all();
    return true;
}
__device__ __forceinline__ void ring_wait(const int* p, int need) {
    while (ld_relaxed_gpu(p) < need) __nanosleep(32);
    (void)ld_acquire_gpu(p);
    fence_proxy_async_all();
}
// after this lane's TMA stores have COMPLETED: publish
__device__ __forceinline__ void ring_signal_stored(int* p) {
    tma_store_wait_all<0>();
    fence_proxy_async_all();
    red_release_gpu(p);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) { return ((uint64_t)kDescHiSw128 << 32) | lo; }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}
template <int N> __device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// dense pair-tile index -> utterance / first cell (binary search in the shared-memory prefix table)
struct PTile {
    int b, first_cell, n_cells, U1b;
};
__device__ __forceinline__ void ptile(int q, const int* s_prefix, int B, PTile& t) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_prefix[mid] <= q) lo = mid; else hi = mid;
    }
    const int tu = s_prefix[B + 1 + lo];
    t.b = lo;
    t.U1b = tu >> 16;
    t.n_cells = (tu & 0xffff) * t.U1b;
    t.first_cell = (q - s_prefix[lo]) * kPairM;
}
// W roles of vocab group g (256-row slabs inside [g kVG, min(V, (g+1) kVG)))
__device__ __forceinline__ int roles_in_group(int g, int V) {
    return (min(V - g * kVG, kVG) + 255) >> 8;
}

// ---- prefix table + flag reset (one block) ------------------------------------------------------
__global__ void __launch_bounds__(kMaxB) ring_prep_kernel(const int* __restrict__ tlen, const int* __restrict__ ulen,
                                                          int B, int T, int U1, int* __restrict__ prefix,
                                                          int* __restrict__ flags, int nflags) {
    __shared__ int s_warp[32];
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    int n = 0, tu = 0;
    if (b < B) {
        const int T_b = min(max(tlen[b], 1), T), U1b = min(max(ulen[b], 0), U1 - 1) + 1;
        n = (T_b * U1b + kPairM - 1) / kPairM;
        tu = T_b | (U1b << 16);
    }
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    if (warp > 0) incl += s_warp[warp - 1];
    if (b == 0) prefix[0] = 0;
    if (b < B) {
        prefix[b + 1] = incl;
        prefix[B + 1 + b] = tu;
    }
    for (int i = b; i < nflags; i += blockDim.x) flags[i] = 0;
}

// =================================================================================================
// P role: recompute z = h W^T tile by tile, emit dz and h into the ring
// =================================================================================================
struct __align__(16) PBars {
    uint64_t b_full[kPBStages], b_empty[kPBStages];
    uint64_t a_full[kMaxKBlocks], a_empty[kMaxKBlocks];
    uint64_t h_ready[kMaxKBlocks];   // local: this CTA's producer warps have written block kb
    uint64_t acc_full[2], acc_empty[2];
    uint64_t slot_free[2];           // local: the ring manager (warp 2) has seen the item's z slot released
    uint64_t stored[2];              // local: every epilogue warp's TMA stores of the item have completed
    uint32_t tmem_base;
    uint32_t pad[3];
};
constexpr int kPEpiThreads = 256;    // warps 4-11
constexpr int kPProdWarps = 8;       // warps 12-19

// per-cell scalars of the dz epilogue (thread == cell)
struct PCell {
    float c2;       // log2|cs| - lse * log2e with cs = g * (gamma_blank + gamma_label): dz = +-exp2(z log2e + c2)
                    // (-1e30 for padding rows and cs == 0: exp2 -> 0)
    uint32_t sgn;   // 0x80008000 if cs < 0 (a negative grad_cost), xor-ed into the packed bf16 pairs
    float dblk;     // dz at the blank column:  cs * p_blank - g * gamma_blank (- g * gamma_label if label == blank)
    float dlab;     // dz at the label column:  cs * p_label - g * gamma_label
    int lab;        // label of the cell's emit transition, -1 if none (or == blank)
};
struct PRaw {
    float g, lse, gl;
    float2 gm, lp;
    int lab, valid;
};
__device__ __forceinline__ void p_load_raw(PRaw& r, const PTile& ti, int m, const RingArgs& a) {
    r.g = 0.f; r.lse = 0.f; r.gl = 0.f; r.gm = make_float2(0.f, 0.f); r.lp = make_float2(0.f, 0.f); r.lab = -1; r.valid = 0;
    if (m < ti.n_cells) {
        const int t = m / ti.U1b, u = m - t * ti.U1b;
        const size_t cell = ((size_t)ti.b * a.T + t) * a.U1 + u;
        r.valid = 1;
        r.lse = __ldg(a.lse + cell);
        if (a.plain) {   // dz = g softmax(z): grad_cost is PER CELL here, unit weight, no blank / label patches
            r.g = __ldg(a.grad_cost + cell);
            r.gm = make_float2(1.f, 0.f);
            return;
        }
        r.g = __ldg(a.grad_cost + ti.b);
        if (a.grad_lse) r.gl = __ldg(a.grad_lse + cell);
        r.gm = __ldg(reinterpret_cast<const float2*>(a.gamma2) + cell);
        r.lp = __ldg(reinterpret_cast<const float2*>(a.lp2) + cell);
        if (u < ti.U1b - 1) r.lab = __ldg(a.labels + (size_t)ti.b * (a.U1 - 1) + u);
    }
}
__device__ __forceinline__ PCell p_finish(const PRaw& r, int V, int blank) {
    PCell s;
    const float cs = fmaf(r.g, r.gm.x + r.gm.y, r.gl);   // coefficient of softmax(z) in dz
    s.c2 = (r.valid && cs != 0.f) ? fmaf(-r.lse, kLog2e, log2f(fabsf(cs))) : -1e30f;
    s.sgn = cs < 0.f ? 0x80008000u : 0u;
    const int lab = r.lab < 0 ? -1 : min(r.lab, V - 1);
    const float cb = r.g * r.gm.x, cl = r.g * r.gm.y;
    s.dblk = r.valid ? fmaf(cs, ex2_approx(r.lp.x * kLog2e), -cb) : 0.f;
    s.dlab = (r.valid && lab >= 0) ? fmaf(cs, ex2_approx(r.lp.y * kLog2e), -cl) : 0.f;
    if (lab == blank) { s.dblk -= cl; s.lab = -1; } else s.lab = lab;
    return s;
}

// One 32-column group of this thread's row: dz = +-exp2(acc log2e + (bias log2e + c2)) -> bf16 into the warp's
// staging buffer ([32 cells x 32 v], 64B swizzle, conflict-free 16-byte stores); the blank / label entries are
// overwritten with their exact fp32 values (from the forward's lp2) before the TMA store.  `bias` holds
// b_out * log2e.  Packed fp32 pairs (FADD2 / FFMA2): 2.5 instructions per element with the MUFU and the bf16 pack.
// (Tried: three of four exponentials as a polynomial on the FMA pipe, FlashAttention-4 style -- the producers' tanh
// got faster but the epilogue, which is bound by issue latency with two warps per scheduler and not by the MUFU,
// got slower: kernel 2.11 -> 2.25 ms.)
// live == false: a 32-column group past the vocabulary (V % 64 == 32): zeros, so that the consumers' 64-wide K
// blocks never see stale ring content.
__device__ __forceinline__ void dz_group(const uint32_t (&r)[32], const float* __restrict__ bias, int v0,
                                         const PCell& s, bool anyneg, int blank, const CUtensorMap* tmap_z,
                                         uint8_t* zbuf, int zcol, int zrow, int lane, bool live
                                         EMO_PROF(, long long& p_rd, long long& p_st)) {
    uint32_t o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0u;
    const float2 c2 = make_float2(s.c2, s.c2), l2 = make_float2(kLog2e, kLog2e);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        if (!live) break;
        const float4 bv = *reinterpret_cast<const float4*>(bias + i);
        const float2 t0 = __ffma2_rn(make_float2(__uint_as_float(r[i + 0]), __uint_as_float(r[i + 1])), l2,
                                     __fadd2_rn(make_float2(bv.x, bv.y), c2));
        const float2 t1 = __ffma2_rn(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), l2,
                                     __fadd2_rn(make_float2(bv.z, bv.w), c2));
        o[i >> 1] = pack_bf16x2(ex2_approx(t0.x), ex2_approx(t0.y));
        o[(i >> 1) + 1] = pack_bf16x2(ex2_approx(t1.x), ex2_approx(t1.y));
    }
    if (anyneg) {   // warp-uniform, false unless some utterance has a negative grad_cost
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] ^= s.sgn;
    }
    EMO_PROF(long long p_c = clock64();)
    if (lane == 0) tma_store_wait_read<kZStBufs - 1>();   // the block before the previous one has left its buffer
    __syncwarp();
    EMO_PROF(p_rd += clock64() - p_c; p_c = clock64();)
    uint8_t* rowp = zbuf + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(rowp + ((i ^ sw) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    const int dl = s.lab - v0;
    if (live && (unsigned)dl < 32u)
        *reinterpret_cast<__nv_bfloat16*>(rowp + (((dl >> 3) ^ sw) << 4) + (dl & 7) * 2) = __float2bfloat16_rn(s.dlab);
    const int db = blank - v0;
    if (live && (unsigned)db < 32u)   // warp-uniform
        *reinterpret_cast<__nv_bfloat16*>(rowp + (((db >> 3) ^ sw) << 4) + (db & 7) * 2) = __float2bfloat16_rn(s.dblk);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(tmap_z, smem_u32(zbuf), zcol, zrow);
        tma_store_commit();
    }
    EMO_PROF(p_st += clock64() - p_c;)
}

// A-operand producer: this warp's 16 rows of the 128-row tile, one 64-wide K block (see joint_bf16.cu)
__device__ __forceinline__ void produce_h16(const uint4 (&re)[4], const uint4 (&rd)[4], int pw, int rsub, int c,
                                            uint8_t* blk, bool plain) {
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

__device__ __forceinline__ void role_produce(const RingArgs& a, const CUtensorMap* tm_w, const CUtensorMap* tm_hst,
                                             const CUtensorMap* tm_zst, uint8_t* smem, int pidx) {
    const int KB = a.J / kBlockK;
    const int NC = (a.V + kChunkN - 1) / kChunkN;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)KB * kABlockBytes;
    uint8_t* sZst = sB + (size_t)kPBStages * kStageBytes;
    PBars* bars = reinterpret_cast<PBars*>(sZst + (kPEpiThreads / 32) * kZStBufs * kZStBytes);
    float* s_bias = reinterpret_cast<float*>(bars + 1);               // [2][kChunkN]
    int* s_prefix = reinterpret_cast<int*>(s_bias + 2 * kChunkN);    // [B+1] + [B]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    int* z_ready = a.flags;
    int* z_done = a.flags + a.NZ;
    int* h_ready_g = a.flags + 2 * a.NZ;
    int* h_done = a.flags + 2 * a.NZ + a.NH;
    const int roles_v = (a.V + 255) >> 8;

    for (int i = threadIdx.x; i < 2 * a.B + 1; i += kRingThreads) s_prefix[i] = __ldg(a.prefix + i);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kPBStages; ++i) {
            mbar_init(smem_u32(&bars->b_full[i]), 2);
            mbar_init(smem_u32(&bars->b_empty[i]), 1);
        }
        for (int i = 0; i < kMaxKBlocks; ++i) {
            mbar_init(smem_u32(&bars->a_full[i]), 2 * kPProdWarps);
            mbar_init(smem_u32(&bars->a_empty[i]), 2);   // MMA commit + h store done
            mbar_init(smem_u32(&bars->h_ready[i]), kPProdWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->acc_full[i]), 1);
            mbar_init(smem_u32(&bars->acc_empty[i]), 2 * (kPEpiThreads / 32));
            mbar_init(smem_u32(&bars->slot_free[i]), 1);
            mbar_init(smem_u32(&bars->stored[i]), kPEpiThreads / 32);
        }
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(tm_w);
        tma_prefetch_desc(tm_hst);
        tma_prefetch_desc(tm_zst);
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // Work unit of a producer pair = ring item = (pair tile q, vocab group g): units are dealt round-robin in item
    // order, so the nP units in production at any time are CONSECUTIVE items and fit the ring whatever V is (a pair
    // that kept a whole tile -- G items -- to itself would need nP * G slots to keep every producer busy).  For
    // G > 1 the h tile is rebuilt per unit (cheap next to 4 vocab chunks of MMAs); G == 1 is one unit per tile.
    constexpr int kCPG = kVG / kChunkN;        // vocab chunks per ring item
    const int NI = s_prefix[a.B] * a.G;

    // register budget per warpgroup (x128 threads): control 40, two epilogue groups 128, two producer groups 88
    // = 472 of the 480 the CTA owns at launch (96 x 640); an exact fit deadlocks in setmaxnreg.inc
    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0) {
            // ===================== TMA producer: w_out tiles [128 v x 64 j] per CTA =====================
            if (lane == 0) {
                uint32_t stage = 0, phase = 0;
                for (int it = pidx; it < NI; it += a.nP) {
                    const int nc0 = (it % a.G) * kCPG, nc1 = min(nc0 + kCPG, NC);
                    for (int nc = nc0; nc < nc1; ++nc) {
                        const int n = min(kChunkN, a.V - nc * kChunkN);
                        const int y = nc * kChunkN + (int)rank * (n >> 1);
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(smem_u32(&bars->b_empty[stage]), phase ^ 1);
                            const uint32_t full = smem_u32(&bars->b_full[stage]);
                            mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), kStageBytes);
                            tma_load_2d_pair(smem_u32(sB + (size_t)stage * kStageBytes), tm_w, kb * kBlockK, y, full);
                            if (++stage == kPBStages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (leader) {
                uint32_t stage = 0, phase = 0, cc = 0, tl = 0;
                const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
                EMO_PROF(long long p_acc = 0, p_a = 0, p_b = 0, p_t0 = clock64(), p_c;)
                for (int it = pidx; it < NI; it += a.nP) {
                    const int nc0 = (it % a.G) * kCPG, nc1 = min(nc0 + kCPG, NC);
                    for (int nc = nc0; nc < nc1; ++nc, ++cc) {
                        const uint32_t buf = cc & 1;
                        EMO_PROF(p_c = clock64();)
                        mbar_wait(smem_u32(&bars->acc_empty[buf]), ((cc >> 1) & 1) ^ 1);
                        EMO_PROF(p_acc += clock64() - p_c;)
                        const int n = min(kChunkN, a.V - nc * kChunkN);
                        const uint32_t idesc = umma_idesc_bf16(kPairM, n);
                        const uint32_t d_tmem = tmem_base + buf * kChunkN;
                        for (int kb = 0; kb < KB; ++kb) {
                            EMO_PROF(p_c = clock64();)
                            if (nc == nc0) mbar_wait(smem_u32(&bars->a_full[kb]), tl & 1);
                            EMO_PROF(p_a += clock64() - p_c; p_c = clock64();)
                            mbar_wait(smem_u32(&bars->b_full[stage]), phase);
                            EMO_PROF(p_b += clock64() - p_c;)
                            tc_fence_after();
                            if (elect_one_sync()) {
                                const uint32_t a_lo = a_lo0 + kb * (kABlockBytes >> 4);
                                const uint32_t b_lo = b_lo0 + stage * (kStageBytes >> 4);
#pragma unroll
                                for (int k16 = 0; k16 < kBlockK / 16; ++k16)
                                    umma_bf16_pair(d_tmem, mk_desc(a_lo + 2 * k16), mk_desc(b_lo + 2 * k16), idesc,
                                                   (kb | k16) != 0);
                                umma_commit_pair(smem_u32(&bars->b_empty[stage]));
                                if (nc == nc1 - 1) umma_commit_pair(smem_u32(&bars->a_empty[kb]));
                                if (kb == KB - 1) umma_commit_pair(smem_u32(&bars->acc_full[buf]));
                            }
                            __syncwarp();
                            if (++stage == kPBStages) { stage = 0; phase ^= 1; }
                        }
                    }
                    ++tl;
                }
                EMO_PROF(if (pidx == 0 && lane == 0)
                             printf("ring P issuer: total %lld clk, %u tiles; wait acc_empty %lld a_full %lld b_full %lld\n",
                                    clock64() - p_t0, tl, p_acc, p_a, p_b);)
            }
        } else if (warp == 2) {
            // ===================== ring manager: slot gate + publication of this CTA's dz items =====================
            // The global-scope release (fence + red.release.gpu, ~800 clk) and the polling of the slot gate used to
            // sit in lane 0 of every epilogue warp: 14 % of the epilogue's time (ncu stall_membar + the poll), on the
            // role's critical path.  Here: the epilogue warps only wait for their own TMA stores and arrive on
            // `stored`; this lane publishes the item (one arrival per CTA) and opens `slot_free` for the next one.
            // Both are tried in turn without blocking (the gate of item k may depend on other pairs' publications,
            // the publication of item k-1 must never wait for it); at most one item is gated ahead.
            if (lane == 0) {
                const int n_it = pidx < NI ? (NI - pidx + a.nP - 1) / a.nP : 0;
                int kg = 0, kp = 0;
                while (kp < n_it) {
                    bool moved = false;
                    if (kg < n_it && kg <= kp + 1) {
                        const int item = pidx + kg * a.nP, zs = item % a.NZ;
                        if (ring_try(z_done + zs, (item / a.NZ) * (1 + roles_in_group(zs % a.G, a.V)))) {
                            mbar_arrive(smem_u32(&bars->slot_free[kg & 1]));
                            ++kg;
                            moved = true;
                        }
                    }
                    if (kp < kg && mbar_try_wait(smem_u32(&bars->stored[kp & 1]), (kp >> 1) & 1)) {
                        fence_proxy_async_all();
                        red_release_gpu(z_ready + (pidx + kp * a.nP) % a.NZ);
                        ++kp;
                        moved = true;
                    }
                    if (!moved) __nanosleep(32);
                }
            }
        } else if (warp == 3) {
            // ===================== h writer: every finished h block -> ring =====================
            if (lane == 0) {
                uint32_t tl = 0;
                for (int it = pidx; it < NI; it += a.nP, ++tl) {
                    const int q = it / a.G;
                    if (it % a.G != 0) {   // the tile's h goes to the ring with its first vocab group only
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(smem_u32(&bars->h_ready[kb]), tl & 1);
                            mbar_arrive(smem_u32(&bars->a_empty[kb]));
                        }
                        continue;
                    }
                    const int hs = q % a.NH;
                    ring_wait(h_done + hs, (q / a.NH) * roles_v);
                    const int row0 = hs * kPairM + (int)rank * kTileM;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(smem_u32(&bars->h_ready[kb]), tl & 1);
                        tma_store_2d(tm_hst, smem_u32(sA + (size_t)kb * kABlockBytes), kb * kBlockK, row0);
                        tma_store_commit();
                        tma_store_wait_read<0>();
                        mbar_arrive(smem_u32(&bars->a_empty[kb]));
                    }
                    ring_signal_stored(h_ready_g + hs);
                }
            }
        }
    } else if (warp < 12) {
        reg_inc<128>();
        // ===================== epilogue: z -> dz -> ring =====================
        // Two warps per TMEM lane quadrant: warp (qd, hf) owns columns [128 hf, 128 hf + 128) of every 256-wide
        // chunk for the rows of quadrant qd; thread == lattice cell.
        const int qd = warp & 3, hf = (warp - 4) >> 2;
        const int row = qd * 32 + lane;
        const int etid = threadIdx.x - 128;   // 0..255
        uint8_t* zbuf0 = sZst + (warp - 4) * kZStBufs * kZStBytes;   // alternating blocks (g & 1)
        const uint32_t acc_empty0 = mapa_shared(smem_u32(&bars->acc_empty[0]), 0);
        uint32_t cc = 0, kit = 0;              // kit: index of the item in this pair's sequence
        const int pblank = a.plain ? -1 : a.blank;   // plain: no blank patch
        int pending = -1;                      // z slot whose stores have been issued but not yet handed to the manager
        EMO_PROF(long long p_sig = 0, p_gate = 0, p_accw = 0, p_t0 = clock64(), p_rd = 0, p_st = 0, p_bar = 0, p_ldw = 0;)
        float nb;                              // bias * log2e of the chunk about to be processed (see dz_group)
        PTile ti;
        PRaw nxt;
        if (pidx < NI) {
            ptile(pidx / a.G, s_prefix, a.B, ti);
            p_load_raw(nxt, ti, ti.first_cell + (int)rank * kTileM + row, a);
        }
        {   // bias of the first chunk of this pair's first unit
            const int i0 = (pidx % a.G) * kCPG * kChunkN + etid;
            nb = i0 < a.V ? __ldg(a.b_out + i0) * kLog2e : 0.f;
        }
        for (int it = pidx; it < NI; it += a.nP) {
            const int nc0 = (it % a.G) * kCPG, nc1 = min(nc0 + kCPG, NC);
            const PCell cur = p_finish(nxt, a.V, a.blank);
            const bool anyneg = __any_sync(0xffffffffu, cur.sgn != 0u);
            if (it + a.nP < NI) {   // next unit's scalars stay in flight during this one
                ptile((it + a.nP) / a.G, s_prefix, a.B, ti);
                p_load_raw(nxt, ti, ti.first_cell + (int)rank * kTileM + row, a);
            }
            const int nnc0 = ((it + a.nP) % a.G) * kCPG;   // first chunk of the next unit
            for (int nc = nc0; nc < nc1; ++nc, ++cc) {
                const uint32_t buf = cc & 1;
                const int n = min(kChunkN, a.V - nc * kChunkN);
                float* bias = s_bias + buf * kChunkN;
                bias[etid] = nb;
                {   // prefetch the next chunk's bias (the first chunk of the next unit after the last one of this)
                    const int nn = (nc + 1 == nc1) ? nnc0 : nc + 1;
                    const int i0 = nn * kChunkN + etid;
                    nb = i0 < a.V ? __ldg(a.b_out + i0) * kLog2e : 0.f;
                }
                EMO_PROF(const long long p_cb = clock64();)
                named_bar_sync(1, kPEpiThreads);
                EMO_PROF(const long long p_c0 = clock64(); p_bar += p_c0 - p_cb;)
                mbar_wait(smem_u32(&bars->acc_full[buf]), (cc >> 1) & 1);
                EMO_PROF(p_accw += clock64() - p_c0;)
                tc_fence_after();
                const int item = it;               // == q * G + nc / kCPG
                const int zs = item % a.NZ;
                if (nc == nc0) {
                    // first chunk of a ring item: hand the previous item to the ring manager (its stores were issued
                    // a chunk ago), then wait until the manager has seen this slot's previous content released
                    if (lane == 0) {
                        EMO_PROF(const long long p_c = clock64();)
                        if (pending >= 0) {
                            tma_store_wait_all<0>();
                            mbar_arrive(smem_u32(&bars->stored[(kit - 1) & 1]));
                        }
                        EMO_PROF(p_sig += clock64() - p_c;)
                        mbar_wait(smem_u32(&bars->slot_free[kit & 1]), (kit >> 1) & 1);
                        EMO_PROF(p_gate += clock64() - p_c;)
                    }
                    pending = zs;
                    ++kit;
                    __syncwarp();
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + buf * kChunkN;
                const int gl = n >> 5;                                  // live 32-column groups of the chunk
                const int g0 = hf * 4, g1 = min(hf * 4 + 4, (gl + 1) & ~1);   // mine (padded to whole 64-wide K blocks)
                const int zrow = zs * kPairM + (int)rank * kTileM + qd * 32;
                const int zcol0 = (nc & 3) * kChunkN;
                if (g0 < g1) {
                    uint32_t ra[32], rb[32];
                    tmem_ld_32x32b_x32(taddr + g0 * 32, ra);
                    for (int g = g0; g < g1; g += 2) {
                        EMO_PROF(const long long p_cl = clock64();)
                        tmem_wait_ld();
                        EMO_PROF(p_ldw += clock64() - p_cl;)
                        if (g + 1 < g1) tmem_ld_32x32b_x32(taddr + (g + 1) * 32, rb);
                        dz_group(ra, bias + g * 32, nc * kChunkN + g * 32, cur, anyneg, pblank, tm_zst, zbuf0,
                                 zcol0 + g * 32, zrow, lane, g < gl EMO_PROF(, p_rd, p_st));
                        if (g + 1 < g1) {
                            tmem_wait_ld();
                            if (g + 2 < g1) tmem_ld_32x32b_x32(taddr + (g + 2) * 32, ra);
                            dz_group(rb, bias + (g + 1) * 32, nc * kChunkN + (g + 1) * 32, cur, anyneg, pblank, tm_zst,
                                     zbuf0 + kZStBytes, zcol0 + (g + 1) * 32, zrow, lane, g + 1 < gl EMO_PROF(, p_rd, p_st));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty0 + buf * 8);
            }
        }
        if (lane == 0) {
            tma_store_wait_all<0>();
            if (pending >= 0) mbar_arrive(smem_u32(&bars->stored[(kit - 1) & 1]));
        }
        EMO_PROF(if (pidx == 0 && warp == 4 && lane == 0)
                     printf("ring P epilogue warp 4: total %lld clk; wait acc_full %lld, publish %lld, publish + slot gate %lld, "
                            "named barrier %lld, tmem ld wait (even groups) %lld, staging wait_read %lld, sts+fence+tma issue %lld\n",
                            clock64() - p_t0, p_accw, p_sig, p_gate, p_bar, p_ldw, p_rd, p_st);)
    } else {
        reg_dec<88>();
        // ===================== A producers: h = tanh(enc + dec) -> bf16 -> K-major SW128 smem =====================
        const int pw = warp - 12;      // 8 producer warps, 16 rows each
        const int c = lane & 7;        // 16-byte chunk (8 hidden units) inside the 64-wide K block
        const int rsub = lane >> 3;    // 4 rows per warp pass
        // flattened (tile, K block) sequence with two units of loads in flight
        int lq = pidx - a.nP, lunit = KB;      // lq: unit (tile, vocab group) index
        uint32_t eoff[4], doff[4];
        auto issue = [&](uint4 (&re)[4], uint4 (&rd)[4]) -> bool {
            if (lunit == KB) {
                lq += a.nP;
                if (lq >= NI) { lq = NI; return false; }
                PTile t;
                ptile(lq / a.G, s_prefix, a.B, t);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int row = pw * 16 + p * 4 + rsub;
                    const int m = min(t.first_cell + (int)rank * kTileM + row, t.n_cells - 1);  // clamp padding rows
                    const int tt = m / t.U1b, u = m - tt * t.U1b;
                    eoff[p] = (uint32_t)(((size_t)t.b * a.T + tt) * a.J) + c * 8;
                    doff[p] = (uint32_t)(((size_t)t.b * a.U1 + u) * a.J) + c * 8;
                }
                lunit = 0;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                re[p] = __ldg(reinterpret_cast<const uint4*>(a.enc + eoff[p] + lunit * kBlockK));
                rd[p] = a.plain ? make_uint4(0u, 0u, 0u, 0u)
                                : __ldg(reinterpret_cast<const uint4*>(a.dec + doff[p] + lunit * kBlockK));
            }
            ++lunit;
            return true;
        };
        uint32_t wkb = 0, wtl = 0;
        const uint32_t a_full0 = mapa_shared(smem_u32(&bars->a_full[0]), 0);
        auto work = [&](const uint4 (&re)[4], const uint4 (&rd)[4]) {
            mbar_wait(smem_u32(&bars->a_empty[wkb]), (wtl & 1) ^ 1);
            produce_h16(re, rd, pw, rsub, c, sA + (size_t)wkb * kABlockBytes, a.plain != 0);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&bars->h_ready[wkb]));
                mbar_arrive_cluster(a_full0 + wkb * 8);
            }
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
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// =================================================================================================
// D role: dh[256 cells x J] = dz W, dz from the ring
// =================================================================================================
struct __align__(16) DBars {
    uint64_t dz_full[kDZStages], dz_empty[kDZStages];
    uint64_t op_full[kDOpStages], op_empty[kDOpStages];
    uint64_t acc_full, acc_empty;
    uint32_t tmem_base;
    uint32_t pad[3];
};

__device__ __forceinline__ void role_dh(const RingArgs& a, const CUtensorMap* tm_w, const CUtensorMap* tm_z,
                                        const CUtensorMap* tm_d, uint8_t* smem, int didx) {
    const int NKB = (a.V + kBlockK - 1) / kBlockK;
    const int NMMA = (a.J + 255) / 256;
    const uint32_t op_bytes = (uint32_t)a.J * 64;      // this CTA's half of a [64 v x J] w_out block
    uint8_t* sZ = smem;
    uint8_t* sW = sZ + (size_t)kDZStages * kStageBytes;
    uint8_t* sDst = sW + (size_t)kDOpStages * op_bytes;
    DBars* bars = reinterpret_cast<DBars*>(sDst + kDrainWarps * kDrainBufBytes);
    int* s_prefix = reinterpret_cast<int*>(bars + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    int* z_ready = a.flags;
    int* z_done = a.flags + a.NZ;
    const int tpu = tiles128_per_utt(a.T, a.U1);

    for (int i = threadIdx.x; i < 2 * a.B + 1; i += kRingThreads) s_prefix[i] = __ldg(a.prefix + i);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kDZStages; ++i) {
            mbar_init(smem_u32(&bars->dz_full[i]), 2);
            mbar_init(smem_u32(&bars->dz_empty[i]), 1);
        }
        for (int i = 0; i < kDOpStages; ++i) {
            mbar_init(smem_u32(&bars->op_full[i]), 2);
            mbar_init(smem_u32(&bars->op_empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        mbar_init(smem_u32(&bars->acc_empty), 2 * kDrainWarps);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(tm_w);
        tma_prefetch_desc(tm_z);
        tma_prefetch_desc(tm_d);
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int Q = s_prefix[a.B];

    if (warp == 0) {
        // ===================== TMA: w_out blocks [64 v x J/2] (this CTA's half of every MMA's N range) ==========
        if (lane == 0) {
            uint32_t slot = 0, ph = 0;
            for (int q = didx; q < Q; q += a.nD) {
                for (int kb = 0; kb < NKB; ++kb) {
                    mbar_wait(smem_u32(&bars->op_empty[slot]), ph ^ 1);
                    const uint32_t full = smem_u32(&bars->op_full[slot]);
                    mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), op_bytes);
                    uint32_t dst = smem_u32(sW + (size_t)slot * op_bytes);
                    for (int n = 0; n < NMMA; ++n) {
                        const int half = min(256, a.J - n * 256) >> 1;
                        for (int b = 0; b < half; b += kBlockK) {
                            tma_load_2d_pair(dst, tm_w, n * 256 + (int)rank * half + b, kb * kBlockK, full);
                            dst += kBoxBytes;
                        }
                    }
                    if (++slot == kDOpStages) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== TMA: this CTA's dz blocks [128 cells x 64 v] from the ring =====================
        if (lane == 0) {
            uint32_t zs = 0, zph = 0;
            EMO_PROF(long long p_ring = 0, p_t0 = clock64(), p_c;)
            for (int q = didx; q < Q; q += a.nD) {
                for (int kb = 0; kb < NKB; ++kb) {
                    const int item = q * a.G + kb / (kVG / kBlockK);
                    const int rs = item % a.NZ;
                    EMO_PROF(p_c = clock64();)
                    if (kb % (kVG / kBlockK) == 0) ring_wait(z_ready + rs, (item / a.NZ + 1) * 2);
                    EMO_PROF(p_ring += clock64() - p_c;)
                    mbar_wait(smem_u32(&bars->dz_empty[zs]), zph ^ 1);
                    const uint32_t full = smem_u32(&bars->dz_full[zs]);
                    mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), kStageBytes);
                    tma_load_2d_pair(smem_u32(sZ + (size_t)zs * kStageBytes), tm_z, (kb % (kVG / kBlockK)) * kBlockK,
                                     rs * kPairM + (int)rank * kTileM, full);
                    if (++zs == kDZStages) { zs = 0; zph ^= 1; }
                }
            }
            EMO_PROF(if (didx == 0 && leader)
                         printf("ring D dz loader: total %lld clk; waiting for ring items %lld\n", clock64() - p_t0, p_ring);)
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            uint32_t zs = 0, zph = 0, slot = 0, ph = 0, tl = 0;
            const uint32_t z_lo0 = desc_lo(smem_u32(sZ), 16);
            const uint32_t w_lo0 = desc_lo(smem_u32(sW), kBoxBytes);
            EMO_PROF(long long p_acc = 0, p_dz = 0, p_op = 0, p_t0 = clock64(), p_c;)
            for (int q = didx; q < Q; q += a.nD) {
                EMO_PROF(p_c = clock64();)
                mbar_wait(smem_u32(&bars->acc_empty), (tl & 1) ^ 1);
                EMO_PROF(p_acc += clock64() - p_c;)
                for (int kb = 0; kb < NKB; ++kb) {
                    EMO_PROF(p_c = clock64();)
                    mbar_wait(smem_u32(&bars->dz_full[zs]), zph);
                    EMO_PROF(p_dz += clock64() - p_c; p_c = clock64();)
                    mbar_wait(smem_u32(&bars->op_full[slot]), ph);
                    EMO_PROF(p_op += clock64() - p_c;)
                    tc_fence_after();
                    if (elect_one_sync()) {
                        // the ring item has been read completely once its last K block has landed
                        if ((kb + 1) % (kVG / kBlockK) == 0 || kb == NKB - 1)
                            red_release_gpu(z_done + (q * a.G + kb / (kVG / kBlockK)) % a.NZ);
                        const uint32_t a_lo = z_lo0 + zs * (kStageBytes >> 4);
                        const uint32_t b_lo = w_lo0 + slot * (op_bytes >> 4);
#pragma unroll
                        for (int k16 = 0; k16 < kBlockK / 16; ++k16) {
                            for (int n = 0; n < NMMA; ++n) {
                                const int Nn = min(256, a.J - n * 256);
                                umma_bf16_pair(tmem_base + n * 256, mk_desc(a_lo + 2 * k16),
                                               mk_desc(b_lo + n * (2 * kBoxBytes >> 4) + k16 * (2048 >> 4)),
                                               umma_idesc_bf16(kPairM, Nn, 0, 1), (kb | k16) != 0);
                            }
                        }
                        umma_commit_pair(smem_u32(&bars->dz_empty[zs]));
                        umma_commit_pair(smem_u32(&bars->op_empty[slot]));
                        if (kb == NKB - 1) umma_commit_pair(smem_u32(&bars->acc_full));
                    }
                    __syncwarp();
                    if (++zs == kDZStages) { zs = 0; zph ^= 1; }
                    if (++slot == kDOpStages) { slot = 0; ph ^= 1; }
                }
                ++tl;
            }
            EMO_PROF(if (didx == 0 && lane == 0)
                         printf("ring D issuer: total %lld clk, %u tiles; wait acc_empty %lld dz_full %lld op_full %lld\n",
                                clock64() - p_t0, tl, p_acc, p_dz, p_op);)
        }
    } else if (warp >= 20 - kDrainWarps) {
        // ===================== drain: dh -> bf16, tile-major (rows of the valid cells, J) =====================
        // 16 warps: TMEM lane quadrant qd = warp & 3, column quarter cq.  Each warp moves its [32 cells x 32 j]
        // blocks through a private shared-memory buffer (64B swizzle, conflict-free 16-byte stores) and one TMA
        // store per block; the factor (1 - h^2) is applied by the reduction kernel.
        const int dw = warp - (20 - kDrainWarps);
        const int qd = warp & 3, cq = dw >> 2;
        const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
        const uint32_t acc_empty_addr = mapa_shared(smem_u32(&bars->acc_empty), 0);
        const int G = a.J >> 7;              // 32-column groups per column quarter
        const int col_base = cq * (a.J >> 2);
        uint8_t* buf = sDst + dw * kDrainBufBytes;
        uint8_t* rowp = buf + lane * 64;
        const int sw = (lane >> 1) & 3;
        uint32_t tl = 0;
        PTile ti;
        for (int q = didx; q < Q; q += a.nD) {
            ptile(q, s_prefix, a.B, ti);
            const int row0 = (ti.b * tpu + (ti.first_cell + (int)rank * kTileM) / kTileM) * kTileM + qd * 32;
            mbar_wait(smem_u32(&bars->acc_full), tl & 1);
            tc_fence_after();
            auto put = [&](const uint32_t (&x)[32], int g) {
                if (lane == 0) tma_store_wait_read<0>();   // the previous block has left the buffer
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(
                        pack_bf16x2(__uint_as_float(x[8 * j]), __uint_as_float(x[8 * j + 1])),
                        pack_bf16x2(__uint_as_float(x[8 * j + 2]), __uint_as_float(x[8 * j + 3])),
                        pack_bf16x2(__uint_as_float(x[8 * j + 4]), __uint_as_float(x[8 * j + 5])),
                        pack_bf16x2(__uint_as_float(x[8 * j + 6]), __uint_as_float(x[8 * j + 7])));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(tm_d, smem_u32(buf), col_base + g * 32, row0);
                    tma_store_commit();
                }
            };
            uint32_t ra[32], rb[32];
            tmem_ld_32x32b_x32(tmem_base + lane_base + col_base, ra);
            for (int g = 0; g < G; g += 2) {
                tmem_wait_ld();
                if (g + 1 < G) tmem_ld_32x32b_x32(tmem_base + lane_base + col_base + (g + 1) * 32, rb);
                put(ra, g);
                if (g + 1 < G) {
                    tmem_wait_ld();
                    if (g + 2 < G) tmem_ld_32x32b_x32(tmem_base + lane_base + col_base + (g + 2) * 32, ra);
                    put(rb, g + 1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_addr);
            ++tl;
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// =================================================================================================
// W role: dW[256 v x J] += dz^T h over the tiles of this pair's split; d_b_out = column sums of dz
// =================================================================================================
struct __align__(16) WBars {
    uint64_t z_full[kWZStages];     // local: TMA bytes of this CTA's dz blocks
    uint64_t dz_full[kWZStages];    // leader: column-sum warps of both CTAs have read the stage
    uint64_t dz_empty[kWZStages];   // both CTAs (multicast commit): the MMAs have read the stage
    uint64_t op_full[kWOpStages], op_empty[kWOpStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad[1];
};

__device__ __forceinline__ void role_dw(const RingArgs& a, const CUtensorMap* tm_z, const CUtensorMap* tm_h,
                                        uint8_t* smem, int widx) {
    const int NMMA = (a.J + 255) / 256;
    const uint32_t op_bytes = (uint32_t)a.J * 64;      // this CTA's half of a [64 cells x J] h block
    uint8_t* sZ = smem;
    uint8_t* sH = sZ + (size_t)kWZStages * kStageBytes;
    WBars* bars = reinterpret_cast<WBars*>(sH + (size_t)kWOpStages * op_bytes);
    int* s_prefix = reinterpret_cast<int*>(bars + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    int* z_ready = a.flags;
    int* z_done = a.flags + a.NZ;
    int* h_ready_g = a.flags + 2 * a.NZ;
    int* h_done = a.flags + 2 * a.NZ + a.NH;
    const int roles_v = (a.V + 255) >> 8;
    const int role = widx % roles_v, split = widx / roles_v;
    const int gr = (role * 256) / kVG;                       // vocab group of this role's slab
    const int v0_cta = role * 256 + (int)rank * kTileM;      // first vocab row of this CTA
    const int vloc = v0_cta - gr * kVG;                      // column inside the ring item

    for (int i = threadIdx.x; i < 2 * a.B + 1; i += kRingThreads) s_prefix[i] = __ldg(a.prefix + i);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kWZStages; ++i) {
            mbar_init(smem_u32(&bars->z_full[i]), 1);
            mbar_init(smem_u32(&bars->dz_full[i]), 2 * kColsumWarps);
            mbar_init(smem_u32(&bars->dz_empty[i]), 1);
        }
        for (int i = 0; i < kWOpStages; ++i) {
            mbar_init(smem_u32(&bars->op_full[i]), 2);
            mbar_init(smem_u32(&bars->op_empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(tm_z);
        tma_prefetch_desc(tm_h);
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int Q = s_prefix[a.B];
    const int n_tiles = split < Q ? (Q - split + a.nS - 1) / a.nS : 0;
    const int n_kb = 4 * n_tiles;    // 64-cell K blocks (padding rows of a tile carry dz = 0)

    if (warp == 0) {
        // ===================== TMA: h blocks [64 cells x J/2] from the ring =====================
        if (lane == 0) {
            uint32_t slot = 0, ph = 0;
            for (int q = split; q < Q; q += a.nS) {
                const int hs = q % a.NH;
                ring_wait(h_ready_g + hs, (q / a.NH + 1) * 2);
                for (int kh = 0; kh < 4; ++kh) {
                    mbar_wait(smem_u32(&bars->op_empty[slot]), ph ^ 1);
                    const uint32_t full = smem_u32(&bars->op_full[slot]);
                    mbar_arrive_expect_tx_cluster(mapa_shared(full, 0), op_bytes);
                    uint32_t dst = smem_u32(sH + (size_t)slot * op_bytes);
                    for (int n = 0; n < NMMA; ++n) {
                        const int half = min(256, a.J - n * 256) >> 1;
                        for (int b = 0; b < half; b += kBlockK) {
                            tma_load_2d_pair(dst, tm_h, n * 256 + (int)rank * half + b, hs * kPairM + kh * 64, full);
                            dst += kBoxBytes;
                        }
                    }
                    if (++slot == kWOpStages) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== TMA: dz blocks [64 cells x 128 v] of this CTA's vocab rows =====================
        if (lane == 0) {
            uint32_t zs = 0, zph = 0;
            EMO_PROF(long long p_ring = 0, p_t0 = clock64(), p_c;)
            for (int q = split; q < Q; q += a.nS) {
                const int item = q * a.G + gr;
                const int rs = item % a.NZ;
                EMO_PROF(p_c = clock64();)
                ring_wait(z_ready + rs, (item / a.NZ + 1) * 2);
                EMO_PROF(p_ring += clock64() - p_c;)
                for (int kh = 0; kh < 4; ++kh) {
                    mbar_wait(smem_u32(&bars->dz_empty[zs]), zph ^ 1);
                    const uint32_t full = smem_u32(&bars->z_full[zs]);
                    mbar_arrive_expect_tx(full, kStageBytes);
                    const uint32_t dst = smem_u32(sZ + (size_t)zs * kStageBytes);
                    tma_load_2d(dst, tm_z, vloc, rs * kPairM + kh * 64, full);
                    tma_load_2d(dst + kBoxBytes, tm_z, vloc + kBlockK, rs * kPairM + kh * 64, full);
                    if (++zs == kWZStages) { zs = 0; zph ^= 1; }
                }
            }
            EMO_PROF(if (widx == 0 && leader)
                         printf("ring W dz loader: total %lld clk; waiting for ring items %lld\n", clock64() - p_t0, p_ring);)
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            uint32_t zs = 0, zph = 0, slot = 0, ph = 0;
            const uint32_t z_lo0 = desc_lo(smem_u32(sZ), kBoxBytes);
            const uint32_t h_lo0 = desc_lo(smem_u32(sH), kBoxBytes);
            int q = split;
            EMO_PROF(long long p_dz = 0, p_op = 0, p_t0 = clock64(), p_c;)
            for (int kbi = 0; kbi < n_kb; ++kbi) {
                EMO_PROF(p_c = clock64();)
                mbar_wait(smem_u32(&bars->dz_full[zs]), zph);
                EMO_PROF(p_dz += clock64() - p_c; p_c = clock64();)
                mbar_wait(smem_u32(&bars->op_full[slot]), ph);
                EMO_PROF(p_op += clock64() - p_c;)
                tc_fence_after();
                if (elect_one_sync()) {
                    if ((kbi & 3) == 3) {   // the tile's ring items have been read completely by this pair
                        red_release_gpu(z_done + (q * a.G + gr) % a.NZ);
                        red_release_gpu(h_done + q % a.NH);
                    }
                    const uint32_t a_lo = z_lo0 + zs * (kStageBytes >> 4);
                    const uint32_t b_lo = h_lo0 + slot * (op_bytes >> 4);
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16) {
                        for (int n = 0; n < NMMA; ++n) {
                            const int Nn = min(256, a.J - n * 256);
                            umma_bf16_pair(tmem_base + n * 256, mk_desc(a_lo + k16 * (2048 >> 4)),
                                           mk_desc(b_lo + n * (2 * kBoxBytes >> 4) + k16 * (2048 >> 4)),
                                           umma_idesc_bf16(kPairM, Nn, 1, 1), (kbi | k16) != 0);
                        }
                    }
                    umma_commit_pair(smem_u32(&bars->dz_empty[zs]));
                    umma_commit_pair(smem_u32(&bars->op_empty[slot]));
                }
                __syncwarp();
                if ((kbi & 3) == 3) q += a.nS;
                if (++zs == kWZStages) { zs = 0; zph ^= 1; }
                if (++slot == kWOpStages) { slot = 0; ph ^= 1; }
            }
            if (elect_one_sync()) umma_commit_pair(smem_u32(&bars->acc_full));
            __syncwarp();
            EMO_PROF(if (widx == 0 && lane == 0)
                         printf("ring W issuer: total %lld clk, %d K blocks; wait dz_full %lld op_full %lld\n",
                                clock64() - p_t0, n_kb, p_dz, p_op);)
        }
    } else if (warp >= 4 && warp < 4 + kColsumWarps) {
        // ===================== column sums of dz (d_b_out), then the flush of dW =====================
        // thread = (64-wide vocab box, 16-byte chunk c = 8 vocab entries, 4 consecutive cells); each thread owns
        // an 8-wide vocab strip for the whole kernel
        const int tt = threadIdx.x - 128;
        const int box = tt >> 7, t7 = tt & 127;
        const int c = t7 & 7, r0 = (t7 >> 3) * 4;
        const int vb = v0_cta + box * kBlockK + c * 8;
        const uint32_t dz_full0 = mapa_shared(smem_u32(&bars->dz_full[0]), 0);
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + i;
            off[i] = box * kBoxBytes + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4);
        }
        float colsum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) colsum[e] = 0.f;
        uint32_t zs = 0, zph = 0;
        for (int kbi = 0; kbi < n_kb; ++kbi) {
            const uint8_t* st = sZ + (size_t)zs * kStageBytes;
            mbar_wait(smem_u32(&bars->z_full[zs]), zph);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(st + off[i]);
                colsum[0] += __uint_as_float(v.x << 16); colsum[1] += __uint_as_float(v.x & 0xffff0000u);
                colsum[2] += __uint_as_float(v.y << 16); colsum[3] += __uint_as_float(v.y & 0xffff0000u);
                colsum[4] += __uint_as_float(v.z << 16); colsum[5] += __uint_as_float(v.z & 0xffff0000u);
                colsum[6] += __uint_as_float(v.w << 16); colsum[7] += __uint_as_float(v.w & 0xffff0000u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(dz_full0 + zs * 8);
            if (++zs == kWZStages) { zs = 0; zph ^= 1; }
        }
        // ---- d_b_out: lanes with equal (lane & 7) hold the same vocab strip
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            colsum[e] += __shfl_xor_sync(0xffffffffu, colsum[e], 8);
            colsum[e] += __shfl_xor_sync(0xffffffffu, colsum[e], 16);
        }
        if (lane < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (vb + e < a.Vout && colsum[e] != 0.f) atomicAdd(a.d_b_out + vb + e, colsum[e]);
        }
        // ---- flush dW: TMEM lane = vocab row, columns = hidden units
        const int dw = warp - 4;
        const int qd = warp & 3, hf = dw >> 2;
        const int v = v0_cta + qd * 32 + lane;
        const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
        const int G = a.J >> 6;
        const int col_base = hf * (a.J >> 1);
        mbar_wait(smem_u32(&bars->acc_full), 0);
        tc_fence_after();
        if (n_kb > 0) {
            for (int g = 0; g < G; ++g) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + lane_base + col_base + g * 32, r);
                tmem_wait_ld();
                if (v < a.Vout) {
                    float* dst = a.d_w_out + (size_t)v * a.J + col_base + g * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                   __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                }
            }
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
__global__ void __launch_bounds__(kRingThreads, 1)
joint_bwd_ring_kernel(const __grid_constant__ CUtensorMap tm_w_p,    // w_out bf16 (V,J), box [64 j x 128 v]
                      const __grid_constant__ CUtensorMap tm_w_d,    // w_out bf16 (V,J), box [64 j x 64 v]
                      const __grid_constant__ CUtensorMap tm_h_st,   // ring h (NH*256, J), box [64 j x 128 cells]
                      const __grid_constant__ CUtensorMap tm_h_ld,   // ring h, box [64 j x 64 cells]
                      const __grid_constant__ CUtensorMap tm_z_st,   // ring dz (NZ*256, VGW), box [32 v x 32 cells], 64B swizzle
                      const __grid_constant__ CUtensorMap tm_z_ld_d, // ring dz, box [64 v x 128 cells]
                      const __grid_constant__ CUtensorMap tm_z_ld_w, // ring dz, box [64 v x 64 cells]
                      const __grid_constant__ CUtensorMap tm_dh,     // dh out bf16 (rows,J), box [32 j x 32 cells], 64B swizzle
                      const RingArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int pair = blockIdx.x >> 1;
    if (pair < a.nP) role_produce(a, &tm_w_p, &tm_h_st, &tm_z_st, smem, pair);
    else if (pair < a.nP + a.nD) role_dh(a, &tm_w_d, &tm_z_ld_d, &tm_dh, smem, pair - a.nP);
    else role_dw(a, &tm_z_ld_w, &tm_h_ld, smem, pair - a.nP - a.nD);
}

size_t ring_smem_bytes(int B, int J) {
    const size_t prefix = (size_t)(2 * B + 1) * sizeof(int);
    const size_t p = (size_t)(J / kBlockK) * kABlockBytes + (size_t)kPBStages * kStageBytes +
                     (kPEpiThreads / 32) * kZStBufs * kZStBytes + sizeof(PBars) + 2 * kChunkN * sizeof(float) + prefix;
    const size_t d = (size_t)kDZStages * kStageBytes + (size_t)kDOpStages * J * 64 + kDrainWarps * kDrainBufBytes +
                     sizeof(DBars) + prefix;
    const size_t w = (size_t)kWZStages * kStageBytes + (size_t)kWOpStages * J * 64 + sizeof(WBars) + prefix;
    return max(p, max(d, w));
}

struct RingGeom {
    int vgw;            // columns of a z slot
    int G;              // vocab groups per tile
    int NZ, NH;
    size_t z_bytes, h_bytes, flag_bytes, prefix_bytes;
};
RingGeom ring_geom(int B, int J, int V) {
    RingGeom g;
    int slots = kRingSlots;
#if defined(EMO_TUNING) || defined(EMO_ZC_PROF)
    if (const char* e = getenv("EMO_RING_SLOTS")) slots = max(8, atoi(e));   // tuning builds only (tools/)
#endif
    g.vgw = (min(V, kVG) + 63) / 64 * 64;
    g.G = ceil_div(V, kVG);
    g.NZ = slots / g.G * g.G;      // a multiple of G: a slot always serves the same vocab group
    g.NH = slots;
    g.z_bytes = align_up((size_t)g.NZ * kPairM * g.vgw * 2, 1024);
    g.h_bytes = align_up((size_t)g.NH * kPairM * J * 2, 1024);
    g.flag_bytes = align_up((size_t)(2 * g.NZ + 2 * g.NH) * sizeof(int), 256);
    g.prefix_bytes = align_up((size_t)(2 * B + 1) * sizeof(int), 256);
    return g;
}

// resident CTA pairs of the ring kernel on the current device (cached per device)
int ring_max_pairs(size_t smem) {
    static thread_local int cached_dev = -1, cached = 0;
    static thread_local size_t cached_smem = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev == cached_dev && smem == cached_smem) return cached;
    if (cudaFuncSetAttribute(joint_bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count() / 2 * 2);
    cfg.blockDim = dim3(kRingThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, joint_bwd_ring_kernel, &cfg) != cudaSuccess) n = 0;
    cached = min(n, sm_count() / 2);
    cached_dev = dev;
    cached_smem = smem;
    return cached;
}

}  // namespace

bool joint_ring_supported(int B, int T, int U1, int J, int V) {
    return B <= kMaxB && T < 65536 && U1 < 65536 && J % 128 == 0 && J <= kMaxKBlocks * kBlockK && V > 0 &&
           ring_smem_bytes(B, J) <= (size_t)kSmemLimit;
}

size_t joint_ring_workspace(int B, int T, int U1, int J, int V) {
    const RingGeom g = ring_geom(B, J, padded_vocab(V));
    return g.z_bytes + g.h_bytes + g.flag_bytes + g.prefix_bytes;
}

// Role split of the resident pairs.  The three roles execute one GEMM unit each; the producer also does the
// tanh / exp work, whose cost per logit does not depend on J, so its share grows when J shrinks.  Weights measured
// at J = 512 (29 / 21 / 24 of 74 pairs at V = 1024; tools/gpu_ringprof.sh), the two gradient GEMMs scaled by J / 512.
void ring_split(int pairs, int V, int J, int& nP, int& nD, int& nS) {
    const int roles_v = ceil_div(V, 256);
    const float j = (float)J / 512.f;
    const float wP = 0.392f, wD = 0.284f * j, wW = 0.324f * j;
    // W pairs come in multiples of roles_v (every split needs all vocab roles): take the number of splits that
    // minimises the slower of the two sides, W alone or P + D sharing the rest (plain rounding under-provisions W
    // when roles_v is large: V = 4096 has 16 roles, 1.5 splits would be ideal, and 1 leaves W 1.75x behind)
    nS = 1;
    float best = 1e30f;
    for (int c = 1; pairs - c * roles_v >= 2; ++c) {
        const float t = fmaxf(wW / (c * roles_v), (wP + wD) / (pairs - c * roles_v));
        if (t < best) { best = t; nS = c; }
    }
    const int rest = pairs - nS * roles_v;
    nD = min(rest - 1, max(1, (int)(rest * wD / (wP + wD) + 0.5f)));
    nP = rest - nD;               // 74 pairs, V = 1024, J = 512: 29 / 21 / 6 x 4
    // With G > 1 vocab groups per tile the producers take (tile, group) items round-robin: keep nP a multiple of G, so
    // that a pair always serves the same group and the G items of a tile are produced side by side (measured at
    // V = 4096, G = 4: nP = 24 / 28 run 40.9 / 41.1 ms, nP = 25 / 26 / 27 / 29 run 50.4 / 45.2 / 43.5 / 48.7 ms)
    const int G = ceil_div(V, kVG);
    if (G > 1 && nP > G) {
        nP = nP / G * G;
        nD = rest - nP;
    }
#if defined(EMO_TUNING) || defined(EMO_ZC_PROF)
    if (const char* e = getenv("EMO_RING_SPLIT")) {   // "nP,nD,nS": tuning builds only (tools/)
        int p, d, s;
        if (sscanf(e, "%d,%d,%d", &p, &d, &s) == 3 && p > 0 && d > 0 && s > 0 && p + d + s * roles_v <= pairs) {
            nP = p; nD = d; nS = s;
        }
    }
#endif
}

int joint_bwd_ring_launch(const void* w_bf16, const void* enc_h, const void* dec_h, const float* b_out,
                          const int* labels, const int* tlen, const int* ulen, const float* lse, const float* lp2,
                          const float* gamma2, const float* grad_cost, const float* grad_lse, int B, int T, int U1,
                          int J, int V, int Vout, int blank, int plain, void* dh_ws, void* ring_ws, float* d_w_out, float* d_b_out,
                          cudaStream_t st) {
    EMO_REQUIRE(joint_ring_supported(B, T, U1, J, V), EMO_UNSUPPORTED_SHAPE,
                "joint_bwd(bf16, ring): needs B <= %d, J %% 128 == 0, J <= 512", kMaxB);
    const RingGeom g = ring_geom(B, J, V);
    char* zring = (char*)ring_ws;
    char* hring = zring + g.z_bytes;
    int* flags = reinterpret_cast<int*>(hring + g.h_bytes);
    int* prefix = reinterpret_cast<int*>((char*)flags + g.flag_bytes);
    const size_t smem = ring_smem_bytes(B, J);
    const int pairs = ring_max_pairs(smem);
    const int roles_v = ceil_div(V, 256);
    EMO_REQUIRE(pairs >= roles_v + 2, EMO_UNSUPPORTED_SHAPE,
                "joint_bwd(bf16, ring): %d resident CTA pairs cannot host %d vocabulary roles", pairs, roles_v);
    RingArgs a;
    a.enc = (const __half*)enc_h; a.dec = (const __half*)dec_h; a.b_out = b_out; a.labels = labels;
    a.lse = lse; a.lp2 = lp2; a.gamma2 = gamma2; a.grad_cost = grad_cost; a.grad_lse = grad_lse; a.prefix = prefix; a.flags = flags;
    a.d_w_out = d_w_out; a.d_b_out = d_b_out;
    a.B = B; a.T = T; a.U1 = U1; a.J = J; a.V = V; a.Vout = Vout; a.blank = blank; a.plain = plain;
    ring_split(pairs, V, J, a.nP, a.nD, a.nS);
    a.NZ = g.NZ; a.NH = g.NH; a.G = g.G;

    ring_prep_kernel<<<1, kMaxB, 0, st>>>(tlen, ulen, B, T, U1, prefix, flags, 2 * g.NZ + 2 * g.NH);
    EMO_CHECK_LAUNCH("ring_prep_kernel");

    CUtensorMap tm_w_p, tm_w_d, tm_h_st, tm_h_ld, tm_z_st, tm_z_ld_d, tm_z_ld_w, tm_dh;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tm_w_p, w_bf16, (uint64_t)J, (uint64_t)V, kBlockK, kChunkN / 2))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_w_d, w_bf16, (uint64_t)J, (uint64_t)V, kBlockK, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_h_st, hring, (uint64_t)J, (uint64_t)g.NH * kPairM, kBlockK, kTileM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_h_ld, hring, (uint64_t)J, (uint64_t)g.NH * kPairM, kBlockK, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_z_st, zring, (uint64_t)g.vgw, (uint64_t)g.NZ * kPairM, 32, 32,
                                CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_z_ld_d, zring, (uint64_t)g.vgw, (uint64_t)g.NZ * kPairM, kBlockK, kTileM))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_z_ld_w, zring, (uint64_t)g.vgw, (uint64_t)g.NZ * kPairM, kBlockK, 64))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm_dh, dh_ws, (uint64_t)J, (uint64_t)B * tiles128_per_utt(T, U1) * kTileM, 32, 32,
                                CU_TENSOR_MAP_SWIZZLE_64B))) return rc;

    EMO_CUDA(cudaFuncSetAttribute(joint_bwd_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (a.nP + a.nD + a.nS * roles_v));
    cfg.blockDim = dim3(kRingThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    // Cooperative launch: the roles wait for each other through the ring, so every CTA must be resident at the same
    // time.  The grid is sized to the resident clusters (ring_max_pairs); the attribute makes the driver guarantee it
    // -- the kernel does not start until all of it fits -- also when another stream's kernels hold SMs at that moment.
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    EMO_CUDA(cudaLaunchKernelEx(&cfg, joint_bwd_ring_kernel, tm_w_p, tm_w_d, tm_h_st, tm_h_ld, tm_z_st, tm_z_ld_d,
                                tm_z_ld_w, tm_dh, a));
    EMO_CHECK_LAUNCH("joint_bwd_ring_kernel");
    return EMO_OK;
}


// ---- workspace / launch accounting of the bf16 joint (include/emoasr_b200.h) ---------------------
static size_t casts_bytes(int B, int T, int U1, int J, int V) {
    const size_t Vp = (size_t)padded_vocab(V);
    return align_up(Vp * J * sizeof(__nv_bfloat16), 256) + align_up((size_t)B * T * J * sizeof(__half), 256) +
           align_up((size_t)B * U1 * J * sizeof(__half), 256) + align_up(Vp * sizeof(float), 256);
}

size_t joint_bf16_workspace(int op, int B, int T, int U1, int J, int V) {
    if (op == EMO_OP_RNNT_JOINT_FWD) return casts_bytes(B, T, U1, J, V);
    if (op == EMO_OP_RNNT_JOINT_BWD) {
        // bf16 w_out + fp16 streams, tile-major dh (bf16, rows of the valid cells), the dz / h ring and its flags
        size_t n = casts_bytes(B, T, U1, J, V) + align_up(dh_bytes_for(B, T, U1, J), 1024);
        if (joint_ring_supported(B, T, U1, J, V)) n += joint_ring_workspace(B, T, U1, J, V);
        return n;
    }
    return 0;
}

int joint_bf16_launches(int op, int B, int T, int U1, int J, int V) {
    (void)B; (void)T; (void)U1; (void)J;
    const int pad = padded_vocab(V) != V;        // + vocabulary padding kernel
    if (op == EMO_OP_RNNT_JOINT_BWD) return 6 + pad;  // 3 casts, ring prep, ring kernel, axis reductions
    return 4 + pad;                                    // weight cast, 2 stream casts, fused joint forward
}

// casts shared by forward and backward: w_out -> bf16, enc_proj / dec_proj -> fp16 (11-bit mantissa, half the
// gather bytes of fp32); layout of the head of every bf16 workspace
int joint_bf16_casts(const float* enc_proj, const float* dec_proj, const float* w_out, const float* b_out, int B, int T,
                     int U1, int J, int V, void* ws, const void** w_bf16, const void** enc_h, const void** dec_h,
                     const float** b_pad, cudaStream_t st) {
    const int Vp = padded_vocab(V);
    const size_t nw = (size_t)V * J, ne = (size_t)B * T * J, nd = (size_t)B * U1 * J;
    __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(ws);
    __half* e = reinterpret_cast<__half*>((char*)ws + align_up((size_t)Vp * J * sizeof(__nv_bfloat16), 256));
    __half* d = reinterpret_cast<__half*>((char*)e + align_up(ne * sizeof(__half), 256));
    float* bp = reinterpret_cast<float*>((char*)d + align_up(nd * sizeof(__half), 256));
    f32_to_bf16_kernel<<<ceil_div(nw, 4 * 256), 256, 0, st>>>(w_out, w, nw);
    EMO_CHECK_LAUNCH("f32_to_bf16_kernel");
    if (Vp != V) {
        const size_t n_tail = (size_t)(Vp - V) * J;
        pad_vocab_kernel<<<ceil_div(max(n_tail, (size_t)Vp), 256), 256, 0, st>>>(w + nw, n_tail, b_out, bp, V, Vp);
        EMO_CHECK_LAUNCH("pad_vocab_kernel");
        *b_pad = bp;
    } else {
        *b_pad = b_out;
    }
    if (enc_h) {
        f32_to_f16_kernel<<<ceil_div(ne, 4 * 256), 256, 0, st>>>(enc_proj, e, ne);
        f32_to_f16_kernel<<<ceil_div(nd, 4 * 256), 256, 0, st>>>(dec_proj, d, nd);
        EMO_CHECK_LAUNCH("f32_to_f16_kernel");
        *enc_h = e;
        *dec_h = d;
    }
    *w_bf16 = w;
    return EMO_OK;
}

int joint_bwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen,
                   const float* lse, const float* lp2, const float* gamma2, const float* grad_cost,
                   const float* grad_lse, int B, int T, int U1, int J, int V, int blank, float* d_enc_proj,
                   float* d_dec_proj, float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    EMO_REQUIRE(enc_proj && dec_proj && w_out && b_out && labels && tlen && ulen && lse && lp2 && gamma2 &&
                    grad_cost && d_enc_proj && d_dec_proj && d_w_out && d_b_out && ws,
                EMO_BAD_ARG, "joint_bwd(bf16): null pointer");
    int rc = check_bf16_shape(B, T, U1, J, V, blank);
    if (rc) return rc;
    EMO_REQUIRE(ws_bytes >= joint_bf16_workspace(EMO_OP_RNNT_JOINT_BWD, B, T, U1, J, V),
                EMO_WORKSPACE_TOO_SMALL, "joint_bwd(bf16): workspace too small");
    EMO_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)w_out & 15) == 0 && ((uintptr_t)d_w_out & 15) == 0 &&
                    ((uintptr_t)enc_proj & 15) == 0 && ((uintptr_t)dec_proj & 15) == 0,
                EMO_BAD_ARG, "joint_bwd(bf16): pointers must be 16-byte (workspace 256-byte) aligned");
    const size_t nw = (size_t)V * J;
    void* dh_ws = (char*)ws + casts_bytes(B, T, U1, J, V);
    void* ring_ws = (char*)dh_ws + align_up(dh_bytes_for(B, T, U1, J), 1024);
    EMO_CUDA(cudaMemsetAsync(d_w_out, 0, nw * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_b_out, 0, (size_t)V * sizeof(float), st));
    EMO_CUDA(cudaMemsetAsync(d_dec_proj, 0, (size_t)B * U1 * J * sizeof(float), st));
    const void *w_bf16, *enc_h, *dec_h;
    const float* b_pad;
    rc = joint_bf16_casts(enc_proj, dec_proj, w_out, b_out, B, T, U1, J, V, ws, &w_bf16, &enc_h, &dec_h, &b_pad, st);
    if (rc) return rc;
    rc = joint_bwd_ring_launch(w_bf16, enc_h, dec_h, b_pad, labels, tlen, ulen, lse, lp2, gamma2, grad_cost, grad_lse, B, T,
                               U1, J, padded_vocab(V), V, blank, 0, dh_ws, ring_ws, d_w_out, d_b_out, st);
    if (rc) return rc;
    return joint_reduce_dh_launch(dh_ws, enc_proj, dec_proj, tlen, ulen, B, T, U1, J, d_enc_proj, d_dec_proj, st);
}

}  // namespace emo

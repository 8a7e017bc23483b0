// CTC forward-backward over the blank-extended label sequence, fused with the log_softmax that
// precedes it in the reference (asr/modeling/decoders/ctc.py:109-113):
//     nn.CTCLoss(blank, reduction="sum", zero_infinity=True)(logits.transpose(1,0).log_softmax(2), ...)
// The kernels read raw logits (B,T,V) once in the forward (row log-sum-exp) and once in the
// backward (softmax - occupancy), never materialising the (T,B,V) log-prob tensor.
//
//   ctc_row_lse_kernel   HBM-bound: one warp per (b,t) row streams the logits through registers (float4 loads, online
//                        max / sum exp, the (row, slab) sequence of a warp is one software pipeline) and gathers the
//                        2U+1 emission log-probs into the alpha / beta work arrays.
//   ctc_lattice_kernel   grid (B, 2): alpha and beta side by side, one CTA per (utterance, direction), thread s owns
//                        extended state s; serial over t, previous frame in double-buffered shared memory; emissions
//                        prefetched one block of 8 frames ahead, results written 8 frames at a time.
//   ctc_grad_warp_kernel HBM-bound (V % 4 == 0): one warp per row, grad = g softmax(z) streamed with float4 loads /
//                        stores, then the <= 2U+1 entries that carry a posterior are patched with atomics on the lines
//                        just written.  ctc_grad_kernel: the same per CTA with a shared-memory vocabulary accumulator,
//                        for vocabularies that are not a multiple of 4.
//   ctc_gather_kernel    emissions alone (backward called without a forward that staged beta).
// The fused head (Linear + log_softmax + CTC without any (B,T,V) tensor) is ctc_head.cu.
#include "common.cuh"

namespace emo {
namespace {

constexpr int kPrefetch = 8;
constexpr int kRowThreads = 256;

__device__ __forceinline__ float block_reduce_lse(float m, float s, float* sm_m, float* sm_s) {
    // warp then block reduce of (max, sumexp); returns lse to all threads
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm_m[warp] = m; sm_s[warp] = s; }
    __syncthreads();
    int nw = blockDim.x >> 5;
    m = lane < nw ? sm_m[lane] : kNegInf;
    s = lane < nw ? sm_s[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        float s2 = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(m, s, m2, s2);
    }
    __syncthreads();
    return m + logf(s);
}

struct Ext {
    int label;   // l'_s
    bool skip;   // transition s-2 -> s allowed
};

__device__ __forceinline__ Ext ext_state(const long long* __restrict__ y, int s, int S_b, int blank,
                                         int V) {
    Ext e;
    e.label = blank;
    e.skip = false;
    if (s < S_b && (s & 1)) {
        long long l = y[s >> 1];
        l = l < 0 ? 0 : (l >= V ? V - 1 : l);
        e.label = (int)l;
        if (s >= 2) e.skip = e.label != blank && l != y[(s >> 1) - 1];
    }
    return e;
}

constexpr int kRowVec = 8;   // float4 per lane in flight (one 4 KiB slab of the row per warp pass)
constexpr int kLseWarps = 8;

// One WARP per row r = b*T + t (grid-stride), no block-level synchronisation: the row streams through
// registers in slabs of 32 lanes x kRowVec float4 with an online (max, sum exp); then the warp
// gathers the S_b = 2 U_b + 1 emission log-probs lp_ext[r][s] = z[l'_s] - lse (the row is L2-hot) and
// writes them where the alpha (and beta) recursion will find them -- the lattice kernels never
// touch the logits.
__global__ void __launch_bounds__(kLseWarps * 32)
ctc_row_lse_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                   const long long* __restrict__ tlen, const long long* __restrict__ ulen, int B, int T,
                   int V, int Umax, int blank, float* __restrict__ lse, float* __restrict__ lp_a,
                   float* __restrict__ lp_b) {
    const int rows = B * T;
    const int S = 2 * Umax + 1;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kLseWarps + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * kLseWarps;
    // end of a row: warp-reduce the online (max, sum exp), write lse, gather the emissions of the
    // blank-extended label sequence (the row is L2-hot)
    auto finish_row = [&](int r, int b, float m, float s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
            const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
            lse_merge(m, s, m2, s2);
        }
        const float l = m + logf(s);
        if (lane == 0) lse[r] = l;
        const float* row = logits + (size_t)r * V;
        const long long U_bl = ulen[b];
        const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
        const int S_b = 2 * U_b + 1;
        const long long* y = labels + (size_t)b * Umax;
        for (int st = lane; st < S_b; st += 32) {
            const float v = __ldg(row + ext_state(y, st, S_b, blank, V).label) - l;
            lp_a[(size_t)r * S + st] = v;
            if (lp_b) lp_b[(size_t)r * S + st] = v;
        }
    };
    if ((V & 3) != 0) {
        // scalar path (vocabulary not a multiple of 4)
        for (int r = warp_global; r < rows; r += warp_stride) {
            const int b = r / T, t = r - b * T;
            if (t >= max(tlen[b], 1LL)) {   // same clamp as the lattice / gradient kernels (T_b >= 1)
                if (lane == 0) lse[r] = 0.f;
                continue;  // warp-uniform
            }
            const float* row = logits + (size_t)r * V;
            float m = kNegInf, s = 0.f;
            for (int i = lane; i < V; i += 32) {
                const float x = __ldg(row + i);
                const float mn = fmaxf(m, x);
                if (mn > kNegInf) {
                    s = s * __expf(m - mn) + __expf(x - mn);
                    m = mn;
                }
            }
            finish_row(r, b, m, s);
        }
        return;
    }
    // Vector path.  The (row, slab) sequence of this warp is one software pipeline: the loads of the NEXT slab --
    // which may be the first slab of the next row -- are issued before the current slab is consumed, so the
    // reduction / emission gather that ends a row runs with a slab of the next row in flight instead of an idle
    // memory pipe.
    const int n4 = V >> 2;
    int nr = warp_global - warp_stride, nbase = n4;   // load cursor (row, first float4 of the slab)
    auto advance = [&]() -> bool {                    // moves the cursor to the next slab; false at the end
        nbase += kRowVec * 32;
        while (nbase >= n4) {
            nr += warp_stride;
            if (nr >= rows) return false;
            const int bb = nr / T;
            if (nr - bb * T >= max(tlen[bb], 1LL)) {  // padded frame: no log-probs needed (T_b clamped to >= 1)
                if (lane == 0) lse[nr] = 0.f;
                continue;
            }
            nbase = 0;
        }
        return true;
    };
    auto issue = [&](float4 (&x)[kRowVec]) {
        const float4* row4 = reinterpret_cast<const float4*>(logits + (size_t)nr * V);
#pragma unroll
        for (int k = 0; k < kRowVec; ++k) {
            const int i = nbase + k * 32 + lane;
            x[k] = i < n4 ? __ldg(row4 + i) : make_float4(kNegInf, kNegInf, kNegInf, kNegInf);
        }
    };
    float4 xa[kRowVec], xb[kRowVec];
    float m = kNegInf, s = 0.f;
    auto consume = [&](const float4 (&x)[kRowVec], int r, int base) {
        float mx = kNegInf;
#pragma unroll
        for (int k = 0; k < kRowVec; ++k)
            mx = fmaxf(mx, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
        const float mn = fmaxf(m, mx);
        if (mn > kNegInf) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < kRowVec; ++k)
                acc += __expf(x[k].x - mn) + __expf(x[k].y - mn) + __expf(x[k].z - mn) + __expf(x[k].w - mn);
            s = s * __expf(m - mn) + acc;
            m = mn;
        }
        if (base + kRowVec * 32 >= n4) {   // last slab of the row
            finish_row(r, r / T, m, s);
            m = kNegInf;
            s = 0.f;
        }
    };
    bool va = advance();
    int ra = nr, ba = nbase;
    if (va) issue(xa);
    while (va) {
        const bool vb = advance();
        const int rb = nr, bb = nbase;
        if (vb) issue(xb);
        consume(xa, ra, ba);
        if (!vb) break;
        va = advance();
        ra = nr; ba = nbase;
        if (va) issue(xa);
        consume(xb, rb, bb);
    }
}

// grid (B, 2): y = 0 alpha, y = 1 beta; blockDim.x = S rounded up to a warp.  Thread s owns extended
// state s.  The previous frame lives in a double-buffered shared array padded by -inf guards, so the
// three-way recursion is three shared loads and one barrier per frame.  Emissions lp_ext[t][s] are
// read from the slot the result is written to (prefetched kPrefetch frames ahead, so the read of a
// slot always precedes its overwrite).  alpha_t and beta_t both include the emission at t.
__global__ void __launch_bounds__(1024, 1)
ctc_lattice_kernel(const long long* __restrict__ labels, const long long* __restrict__ tlen,
                   const long long* __restrict__ ulen, int T, int V, int Umax, int blank,
                   int zero_infinity, int first_dir, float* __restrict__ alpha_ws,
                   float* __restrict__ beta_ws, float* __restrict__ nll) {
    __shared__ float buf[2][1024 + 4];
    const int b = blockIdx.x;
    const bool backward = (int)blockIdx.y + first_dir == 1;
    const int S = 2 * Umax + 1;
    const int s = threadIdx.x;
    long long T_bl = tlen[b], U_bl = ulen[b];
    const int T_b = (int)(T_bl < 1 ? 1 : (T_bl > T ? T : T_bl));
    const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
    const int S_b = 2 * U_b + 1;
    const long long* y = labels + (size_t)b * Umax;
    const bool jump = backward ? ext_state(y, s + 2, S_b, blank, V).skip : ext_state(y, s, S_b, blank, V).skip;
    const bool valid = s < S_b;
    float* io = (backward ? beta_ws : alpha_ws) + (size_t)b * T * S;
    const int nb = backward ? 1 : -1;  // neighbour direction in state space

    for (int i = threadIdx.x; i < 2 * (1024 + 4); i += blockDim.x) (&buf[0][0])[i] = kNegInf;
    __syncthreads();

    auto frame = [&](int i) { return backward ? T_b - 1 - i : i; };
    // always a valid address (clamped), no select on the loaded value: the load must stay in flight
    // for kPrefetch frames, a select right behind it would stall the thread until it lands
    const int s_ld = valid ? s : 0;
    auto load = [&](int i) -> float { return io[(size_t)frame(min(i, T_b - 1)) * S + s_ld]; };
    // Emissions of the NEXT block of kPrefetch frames are loaded at the top of the current block (kPrefetch
    // independent loads back to back), so every load has a whole block of frames to land.  (A per-frame ring
    // "load frame i + kPrefetch while computing frame i" was scheduled by the compiler with an effective distance
    // of one frame: 63 % of the kernel's stall samples sat on the load's scoreboard.)
    float cur[kPrefetch], nxt[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) cur[k] = load(k);

    // results are kept in registers and written kPrefetch frames at a time: a global store in front of
    // every per-frame barrier would put one store round trip on the serial chain
    float vals[kPrefetch];
    for (int i0 = 0; i0 < T_b; i0 += kPrefetch) {
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) nxt[k] = load(i0 + kPrefetch + k);
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int i = i0 + k;
            if (i >= T_b) break;  // uniform
            const float lp = cur[k];
            const float* prev = buf[(i + 1) & 1] + 2;  // previous frame, index by state
            float val = kNegInf;
            if (valid) {
                if (i == 0) {
                    // forward: alpha_0(0), alpha_0(1); backward: beta_{T-1}(S-1), beta_{T-1}(S-2)
                    const bool init = backward ? (s >= S_b - 2) : (s < 2);
                    val = init ? lp : kNegInf;
                } else {
                    const float a = prev[s];
                    const float n1 = prev[s + nb];
                    const float n2 = jump ? prev[s + 2 * nb] : kNegInf;
                    val = log_add_exp3_fast(a, n1, n2) + lp;
                }
            }
            vals[k] = val;
            buf[i & 1][2 + s] = val;
            __syncthreads();
        }
        if (valid) {
#pragma unroll
            for (int k = 0; k < kPrefetch; ++k)
                if (i0 + k < T_b) io[(size_t)frame(i0 + k) * S + s] = vals[k];
        }
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) cur[k] = nxt[k];
    }
    if (!backward && s == 0) {
        const float* last = buf[(T_b - 1) & 1] + 2;
        float l = log_add_exp(last[S_b - 1], S_b > 1 ? last[S_b - 2] : kNegInf);
        float v = -l;
        if (!(l > kNegInf) || l != l) v = zero_infinity ? 0.f : INFINITY;
        nll[b] = v;
    }
}

// Persistent CTAs, grid-stride over rows (b,t).  acc[V] in dynamic shared memory stays zero
// between rows: only the touched entries are cleared again.  The state posteriors
// occ_t(s) = exp(alpha_t(s) + beta_t(s) - lp_t(l'_s) - ll) are formed here from the two lattices.
__global__ void __launch_bounds__(kRowThreads)
ctc_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                const long long* __restrict__ tlen, const long long* __restrict__ ulen,
                const float* __restrict__ lse, const float* __restrict__ alpha_ws,
                const float* __restrict__ beta_ws, const float* __restrict__ grad_nll, int B, int T,
                int V, int Umax, int blank, float* __restrict__ grad) {
    extern __shared__ float acc[];  // V floats
    const int S = 2 * Umax + 1;
    for (int v = threadIdx.x; v < V; v += kRowThreads) acc[v] = 0.f;
    __syncthreads();
    const int rows = B * T;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        int b = r / T, t = r - b * T;
        long long T_bl = tlen[b], U_bl = ulen[b];
        const int T_b = (int)(T_bl < 1 ? 1 : (T_bl > T ? T : T_bl));
        const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
        const int S_b = 2 * U_b + 1;
        float* grow = grad + (size_t)r * V;
        // feasibility of the utterance from the last alphas (same test as the forward kernel)
        const float* alast = alpha_ws + ((size_t)b * T + (T_b - 1)) * S;
        const float l = log_add_exp(alast[S_b - 1], S_b > 1 ? alast[S_b - 2] : kNegInf);
        const bool feasible = l > kNegInf && l == l && l < INFINITY;
        if (t >= T_b || !feasible) {  // uniform over the CTA
            if ((V & 3) == 0) {
                float4* g4 = reinterpret_cast<float4*>(grow);
                for (int i = threadIdx.x; i < (V >> 2); i += kRowThreads)
                    g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int i = threadIdx.x; i < V; i += kRowThreads) grow[i] = 0.f;
            }
            continue;
        }
        const long long* y = labels + (size_t)b * Umax;
        const float row_lse = lse[r];
        const float* row = logits + (size_t)r * V;
        const float* al = alpha_ws + (size_t)r * S;
        const float* be = beta_ws + (size_t)r * S;
        for (int s = threadIdx.x; s < S_b; s += kRowThreads) {
            const int lab = ext_state(y, s, S_b, blank, V).label;
            const float a = al[s], bt = be[s];
            if (a > kNegInf && bt > kNegInf) {
                const float lp = __ldg(row + lab) - row_lse;
                atomicAdd(&acc[lab], expf(a + bt - lp - l));
            }
        }
        __syncthreads();
        const float g = grad_nll[b];
        if ((V & 3) == 0) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            float4* g4 = reinterpret_cast<float4*>(grow);
            const float4* a4 = reinterpret_cast<const float4*>(acc);
            for (int i = threadIdx.x; i < (V >> 2); i += kRowThreads) {
                float4 x = __ldg(row4 + i);
                float4 a = a4[i];
                float4 o;
                o.x = g * (expf(x.x - row_lse) - a.x);
                o.y = g * (expf(x.y - row_lse) - a.y);
                o.z = g * (expf(x.z - row_lse) - a.z);
                o.w = g * (expf(x.w - row_lse) - a.w);
                g4[i] = o;
            }
        } else {
            for (int i = threadIdx.x; i < V; i += kRowThreads)
                grow[i] = g * (expf(__ldg(row + i) - row_lse) - acc[i]);
        }
        __syncthreads();
        for (int s = threadIdx.x; s < S_b; s += kRowThreads)
            acc[ext_state(y, s, S_b, blank, V).label] = 0.f;
        __syncthreads();
    }
}

// Gradient, one WARP per row, no block barriers and no vocabulary-sized accumulator (V % 4 == 0):
//   1. stream the row once:  grad[v] = g * softmax(z)[v]   (float4 loads / stores; the (row, slab) sequence of the
//      warp is one software pipeline, as in ctc_row_lse_kernel: the next slab -- possibly of the next row -- is in
//      flight while the current one is consumed and while a finished row is patched)
//   2. patch the <= S_b entries that carry a posterior:  grad[l'_s] -= g * occ_t(s)   with atomics on the lines
//      just written (L2 hits); the U_b + 1 blank states are summed in the warp first.
// occ_t(s) = exp(alpha_t(s) + beta_t(s) - lp_t(l'_s) - ll) is formed here from the two lattices.
__global__ void __launch_bounds__(kLseWarps * 32)
ctc_grad_warp_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                     const long long* __restrict__ tlen, const long long* __restrict__ ulen,
                     const float* __restrict__ lse, const float* __restrict__ alpha_ws,
                     const float* __restrict__ beta_ws, const float* __restrict__ grad_nll, int B, int T,
                     int V, int Umax, int blank, float* __restrict__ grad) {
    const int rows = B * T;
    const int S = 2 * Umax + 1;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kLseWarps + (threadIdx.x >> 5);
    const int warp_stride = gridDim.x * kLseWarps;
    const int n4 = V >> 2;
    struct Meta { int r; float row_lse, g, l; };
    int nr = warp_global - warp_stride, nbase = n4;   // load cursor (row, first float4 of the slab)
    Meta nmeta = {0, 0.f, 0.f, 0.f};
    auto advance = [&]() -> bool {
        nbase += kRowVec * 32;
        while (nbase >= n4) {
            nr += warp_stride;
            if (nr >= rows) return false;
            const int b = nr / T, t = nr - b * T;
            const long long T_bl = tlen[b], U_bl = ulen[b];
            const int T_b = (int)(T_bl < 1 ? 1 : (T_bl > T ? T : T_bl));
            const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
            const int S_b = 2 * U_b + 1;
            // feasibility of the utterance from the last alphas (same test as the forward kernel)
            const float* alast = alpha_ws + ((size_t)b * T + (T_b - 1)) * S;
            const float l = log_add_exp(alast[S_b - 1], S_b > 1 ? alast[S_b - 2] : kNegInf);
            const bool feasible = l > kNegInf && l == l && l < INFINITY;
            if (t >= T_b || !feasible) {  // warp-uniform: padded frame or infeasible utterance, zero gradient
                float4* g4 = reinterpret_cast<float4*>(grad + (size_t)nr * V);
                for (int i = lane; i < n4; i += 32) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            nmeta.r = nr; nmeta.row_lse = lse[nr]; nmeta.g = grad_nll[b]; nmeta.l = l;
            nbase = 0;
        }
        return true;
    };
    auto issue = [&](float4 (&x)[kRowVec]) {
        const float4* row4 = reinterpret_cast<const float4*>(logits + (size_t)nr * V);
#pragma unroll
        for (int k = 0; k < kRowVec; ++k) {
            const int i = nbase + k * 32 + lane;
            if (i < n4) x[k] = __ldg(row4 + i);
        }
    };
    auto consume = [&](const float4 (&x)[kRowVec], const Meta& mt, int base) {
        float4* g4 = reinterpret_cast<float4*>(grad + (size_t)mt.r * V);
#pragma unroll
        for (int k = 0; k < kRowVec; ++k) {
            const int i = base + k * 32 + lane;
            if (i < n4) {
                float4 o;
                o.x = mt.g * expf(x[k].x - mt.row_lse);
                o.y = mt.g * expf(x[k].y - mt.row_lse);
                o.z = mt.g * expf(x[k].z - mt.row_lse);
                o.w = mt.g * expf(x[k].w - mt.row_lse);
                g4[i] = o;
            }
        }
        if (base + kRowVec * 32 < n4) return;
        // ---- last slab of the row: patch the entries of the blank-extended label sequence
        __syncwarp();   // the row is written (memory ordering among the lanes) before it is patched
        const int b = mt.r / T;
        const long long U_bl = ulen[b];
        const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
        const int S_b = 2 * U_b + 1;
        const long long* y = labels + (size_t)b * Umax;
        const float* row = logits + (size_t)mt.r * V;
        const float* al = alpha_ws + (size_t)mt.r * S;
        const float* be = beta_ws + (size_t)mt.r * S;
        float* grow = grad + (size_t)mt.r * V;
        float blank_occ = 0.f;
        for (int st = lane; st < S_b; st += 32) {
            const int lab = ext_state(y, st, S_b, blank, V).label;
            const float a = al[st], bt = be[st];
            if (a > kNegInf && bt > kNegInf) {
                const float lp = __ldg(row + lab) - mt.row_lse;
                const float occ = expf(a + bt - lp - mt.l);
                if (lab == blank) blank_occ += occ;
                else atomicAdd(grow + lab, -mt.g * occ);
            }
        }
        blank_occ = warp_sum(blank_occ);
        if (lane == 0 && blank_occ != 0.f) atomicAdd(grow + blank, -mt.g * blank_occ);
    };
    float4 xa[kRowVec], xb[kRowVec];
    bool va = advance();
    Meta ma = nmeta;
    int ba = nbase;
    if (va) issue(xa);
    while (va) {
        const bool vb = advance();
        const Meta mb = nmeta;
        const int bb = nbase;
        if (vb) issue(xb);
        consume(xa, ma, ba);
        if (!vb) break;
        va = advance();
        ma = nmeta; ba = nbase;
        if (va) issue(xa);
        consume(xb, mb, bb);
    }
}

// lp_ext gather alone (backward called without a forward that staged beta)
__global__ void ctc_gather_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                  const long long* __restrict__ tlen, const long long* __restrict__ ulen,
                                  const float* __restrict__ lse, int B, int T, int V, int Umax, int blank,
                                  float* __restrict__ lp_out) {
    const int S = 2 * Umax + 1;
    const size_t n = (size_t)B * T * S;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int st = (int)(idx % S);
        const size_t r = idx / S;
        const int b = (int)(r / T), t = (int)(r % T);
        long long U_bl = ulen[b];
        const int U_b = (int)(U_bl < 0 ? 0 : (U_bl > Umax ? Umax : U_bl));
        const int S_b = 2 * U_b + 1;
        if (t < max(tlen[b], 1LL) && st < S_b)
            lp_out[idx] = __ldg(logits + r * V + ext_state(labels + (size_t)b * Umax, st, S_b, blank, V).label) - lse[r];
    }
}

int check_ctc_args(const void* logits, const void* labels, const void* tlen, const void* ulen,
                   int B, int T, int V, int Umax, int blank) {
    EMO_REQUIRE(logits && labels && tlen && ulen, EMO_BAD_ARG, "ctc: null pointer");
    EMO_REQUIRE(B > 0 && T > 0 && V > 0 && Umax >= 0, EMO_BAD_ARG, "ctc: bad sizes");
    EMO_REQUIRE(blank >= 0 && blank < V, EMO_BAD_ARG, "ctc: blank %d outside [0,%d)", blank, V);
    EMO_REQUIRE(2 * Umax + 1 <= 1024, EMO_UNSUPPORTED_SHAPE,
                "ctc: 2*Umax+1 = %d exceeds 1024 extended states", 2 * Umax + 1);
    EMO_REQUIRE((size_t)V * sizeof(float) <= 200 * 1024, EMO_UNSUPPORTED_SHAPE,
                "ctc: vocabulary %d does not fit the shared-memory accumulator", V);
    return EMO_OK;
}

}  // namespace
}  // namespace emo

namespace emo {
// alpha (and, with beta_ws, beta) lattices over emissions already staged in alpha_ws / beta_ws (ctc_head.cu)
int ctc_lattice_launch(const long long* labels, const long long* tlen, const long long* ulen, int B, int T, int V,
                       int Umax, int blank, int zero_infinity, float* alpha_ws, float* beta_ws, float* nll,
                       cudaStream_t st) {
    EMO_REQUIRE(2 * Umax + 1 <= 1024, EMO_UNSUPPORTED_SHAPE, "ctc: 2*Umax+1 = %d exceeds 1024 extended states",
                2 * Umax + 1);
    const int S = 2 * Umax + 1;
    const int threads = (S + 31) / 32 * 32;
    ctc_lattice_kernel<<<dim3(B, beta_ws ? 2 : 1), threads, 0, st>>>(labels, tlen, ulen, T, V, Umax, blank,
                                                                      zero_infinity, 0, alpha_ws, beta_ws, nll);
    EMO_CHECK_LAUNCH("ctc_lattice_kernel");
    return EMO_OK;
}
}  // namespace emo

using namespace emo;

extern "C" int emo_ctc_fwd(const float* logits, const long long* labels, const long long* tlen,
                           const long long* ulen, int B, int T, int V, int Umax, int blank,
                           int zero_infinity, float* lse, float* alpha_ws, float* beta_ws, float* nll,
                           void* stream) {
    int rc = check_ctc_args(logits, labels, tlen, ulen, B, T, V, Umax, blank);
    if (rc) return rc;
    EMO_REQUIRE(lse && alpha_ws && nll, EMO_BAD_ARG, "ctc_fwd: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rows = B * T;
    int grid = min(ceil_div(rows, kLseWarps), sm_count() * 8);
    ctc_row_lse_kernel<<<grid, kLseWarps * 32, 0, st>>>(logits, labels, tlen, ulen, B, T, V, Umax, blank, lse,
                                                     alpha_ws, beta_ws);
    EMO_CHECK_LAUNCH("ctc_row_lse_kernel");
    int S = 2 * Umax + 1;
    int threads = (S + 31) / 32 * 32;
    // alpha and (when the caller will differentiate) beta run side by side
    ctc_lattice_kernel<<<dim3(B, beta_ws ? 2 : 1), threads, 0, st>>>(labels, tlen, ulen, T, V, Umax, blank,
                                                                      zero_infinity, 0, alpha_ws, beta_ws, nll);
    EMO_CHECK_LAUNCH("ctc_lattice_kernel");
    return EMO_OK;
}

extern "C" int emo_ctc_bwd(const float* logits, const long long* labels, const long long* tlen,
                           const long long* ulen, const float* lse, const float* alpha_ws,
                           const float* nll, const float* grad_nll, int B, int T, int V, int Umax,
                           int blank, int zero_infinity, float* beta_ws, int beta_valid,
                           float* grad_logits, void* stream) {
    int rc = check_ctc_args(logits, labels, tlen, ulen, B, T, V, Umax, blank);
    if (rc) return rc;
    EMO_REQUIRE(lse && alpha_ws && nll && grad_nll && beta_ws && grad_logits, EMO_BAD_ARG,
                "ctc_bwd: null pointer");
    (void)zero_infinity;  // infeasible utterances get zero gradient either way (their nll is 0 or
                          // inf; torch yields NaN for inf without zero_infinity, we return 0)
    cudaStream_t st = (cudaStream_t)stream;
    int S = 2 * Umax + 1;
    if (!beta_valid) {  // the forward did not stage beta: gather the emissions and run the beta lattice now
        size_t n = (size_t)B * T * S;
        ctc_gather_kernel<<<(int)min((size_t)sm_count() * 16, (n + 255) / 256), 256, 0, st>>>(
            logits, labels, tlen, ulen, lse, B, T, V, Umax, blank, beta_ws);
        EMO_CHECK_LAUNCH("ctc_gather_kernel");
        int threads = (S + 31) / 32 * 32;
        ctc_lattice_kernel<<<dim3(B, 1), threads, 0, st>>>(labels, tlen, ulen, T, V, Umax, blank, zero_infinity, 1,
                                                           nullptr, beta_ws, nullptr);
        EMO_CHECK_LAUNCH("ctc_lattice_kernel<beta>");
    }
    if ((V & 3) == 0) {   // warp-per-row streaming version
        const int rows = B * T;
        ctc_grad_warp_kernel<<<min(ceil_div(rows, kLseWarps), sm_count() * 8), kLseWarps * 32, 0, st>>>(
            logits, labels, tlen, ulen, lse, alpha_ws, beta_ws, grad_nll, B, T, V, Umax, blank, grad_logits);
        EMO_CHECK_LAUNCH("ctc_grad_warp_kernel");
        return EMO_OK;
    }
    size_t smem = (size_t)V * sizeof(float);
    if (smem > 48 * 1024)
        EMO_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    int rows = B * T;
    int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(200 * 1024) / max(smem, (size_t)1)));
    int grid = min(rows, sm_count() * per_sm);
    ctc_grad_kernel<<<grid, kRowThreads, smem, st>>>(logits, labels, tlen, ulen, lse, alpha_ws,
                                                     beta_ws, grad_nll, B, T, V, Umax, blank,
                                                     grad_logits);
    EMO_CHECK_LAUNCH("ctc_grad_kernel");
    return EMO_OK;
}

// Shared helpers for the emoasr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/emoasr_b200.h"

namespace emo {

// thread-local error text behind emo_last_error_string()
void set_error(const char* fmt, ...);

#define EMO_REQUIRE(cond, status, ...)      \
    do {                                    \
        if (!(cond)) {                      \
            ::emo::set_error(__VA_ARGS__);  \
            return (status);                \
        }                                   \
    } while (0)

#define EMO_CHECK_LAUNCH(what)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ::emo::set_error("%s: %s (%s)", what, cudaGetErrorName(e__),                \
                             cudaGetErrorString(e__));                                  \
            return EMO_LAUNCH_FAILURE;                                                  \
        }                                                                               \
    } while (0)

#define EMO_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            ::emo::set_error("%s: %s (%s)", #call, cudaGetErrorName(e__),               \
                             cudaGetErrorString(e__));                                  \
            return EMO_LAUNCH_FAILURE;                                                  \
        }                                                                               \
    } while (0)

constexpr float kNegInf = -INFINITY;

// log(exp(a)+exp(b)), exact-ish fp32, -inf safe
__device__ __forceinline__ float log_add_exp(float a, float b) {
    float m = fmaxf(a, b);
    if (m == kNegInf) return kNegInf;
    return m + log1pf(expf(-fabsf(a - b)));
}

// ---- single-instruction ex2 / lg2 (MUFU) and branch-free log-sum-exp built on them.  Used inside the
// serial lattice recursions, where the expansions of expf/logf (range fix-ups, branches) are the
// critical path.  Absolute error of a log-add ~1e-7.  -inf safe: the pivot is clamped to -1e30 so
// that (-inf) - pivot stays -inf (no NaN); an all -inf input gives lg2(0) = -inf.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2eF = 1.4426950408889634f;
constexpr float kLn2F = 0.6931471805599453f;
__device__ __forceinline__ float log_add_exp_fast(float a, float b) {
    const float m = fmaxf(fmaxf(a, b), -1e30f);
    const float s = ex2_ftz((a - m) * kLog2eF) + ex2_ftz((b - m) * kLog2eF);
    return fmaf(lg2_ftz(s), kLn2F, m);
}
__device__ __forceinline__ float log_add_exp3_fast(float a, float b, float c) {
    const float m = fmaxf(fmaxf(fmaxf(a, b), c), -1e30f);
    const float s = ex2_ftz((a - m) * kLog2eF) + ex2_ftz((b - m) * kLog2eF) + ex2_ftz((c - m) * kLog2eF);
    return fmaf(lg2_ftz(s), kLn2F, m);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// online (max, sum exp) pair merge
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
    float mn = fmaxf(m, m2);
    if (mn == kNegInf) { m = mn; s = 0.f; return; }
    s = s * expf(m - mn) + s2 * expf(m2 - mn);
    m = mn;
}

// same, with exp2 of log2e-scaled differences (matches the kernels' ex2.approx arithmetic)
__device__ __forceinline__ void lse_merge_exp2(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    if (mn == kNegInf) { m = mn; s = 0.f; return; }
    s = s * exp2f((m - mn) * 1.4426950408889634f) + s2 * exp2f((m2 - mn) * 1.4426950408889634f);
    m = mn;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();  // SMs of the current device (cached per device)

// ---- entry points implemented across translation units ----
int rnnt_lattice_launch(const float* lp2, const int* tlen, const int* ulen, int B, int T, int U1,
                        float* alpha_ws, float* beta_ws, float* cost, float* gamma2,
                        cudaStream_t st);

int joint_fwd_f32(const float* enc_proj, const float* dec_proj, const float* w_out,
                  const float* b_out, const int* labels, const int* tlen, const int* ulen, int B,
                  int T, int U1, int J, int V, int blank, float* lp2, float* lse, void* ws,
                  size_t ws_bytes, cudaStream_t st);
int joint_bwd_f32(const float* enc_proj, const float* dec_proj, const float* w_out,
                  const float* b_out, const int* labels, const int* tlen, const int* ulen,
                  const float* lse, const float* gamma2, const float* grad_cost, const float* grad_lse, int B,
                  int T, int U1, int J, int V, int blank, float* d_enc_proj, float* d_dec_proj,
                  float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t joint_f32_workspace(int op, int B, int T, int U1, int J, int V);
int joint_f32_launches(int op, int B, int T, int U1, int J, int V);

int joint_fwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen, int B,
                   int T, int U1, int J, int V, int blank, float* lp2, float* lse, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int joint_bwd_bf16(const float* enc_proj, const float* dec_proj, const float* w_out,
                   const float* b_out, const int* labels, const int* tlen, const int* ulen,
                   const float* lse, const float* lp2, const float* gamma2, const float* grad_cost,
                   const float* grad_lse, int B, int T, int U1, int J, int V, int blank, float* d_enc_proj,
                   float* d_dec_proj, float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes, cudaStream_t st);
// ring backward (joint_bwd_ring.cu): the default route
bool joint_ring_supported(int B, int T, int U1, int J, int V);
size_t joint_ring_workspace(int B, int T, int U1, int J, int V);
int joint_bwd_ring_launch(const void* w_bf16, const void* enc_h, const void* dec_h, const float* b_out,
                          const int* labels, const int* tlen, const int* ulen, const float* lse, const float* lp2,
                          const float* gamma2, const float* grad_cost, const float* grad_lse, int B, int T, int U1,
                          int J, int V, int Vout, int blank, int plain, void* dh_ws, void* ring_ws, float* d_w_out, float* d_b_out,
                          cudaStream_t st);
int joint_fwd_launch(const void* w_bf16, const void* enc_h, const void* dec_h, const float* b_pad, const int* labels,
                     const int* tlen, const int* ulen, int B, int T, int U1, int J, int Vp, int blank, float* lp2,
                     float* lse, int plain, cudaStream_t st);
int joint_bf16_casts(const float* enc_proj, const float* dec_proj, const float* w_out, const float* b_out, int B, int T,
                     int U1, int J, int V, void* ws, const void** w_bf16, const void** enc_h, const void** dec_h,
                     const float** b_pad, cudaStream_t st);
// axis reductions of dh (joint_reduce.cu)
int joint_reduce_dh_launch(const void* dh_ws, const float* enc_proj, const float* dec_proj, const int* tlen,
                           const int* ulen, int B, int T, int U1, int J, float* d_enc_proj, float* d_dec_proj,
                           cudaStream_t st);
int ctc_lattice_launch(const long long* labels, const long long* tlen, const long long* ulen, int B, int T, int V,
                       int Umax, int blank, int zero_infinity, float* alpha_ws, float* beta_ws, float* nll,
                       cudaStream_t st);
int joint_reduce_dh_launch_f16(const void* dh_ws, const void* enc16, const void* dec16, const int* tlen, const int* ulen,
                               int B, int T, int U1, int J, float* d_enc_proj, float* d_dec_proj, void* d_enc_bf,
                               cudaStream_t st);
size_t joint_bf16_workspace(int op, int B, int T, int U1, int J, int V);
int joint_bf16_launches(int op, int B, int T, int U1, int J, int V);

}  // namespace emo

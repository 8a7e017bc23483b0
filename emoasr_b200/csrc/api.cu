// C-ABI plumbing: error string, device info, precision dispatch (include/emoasr_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace emo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace emo

using namespace emo;

extern "C" int emo_abi_version(void) { return EMO_ABI_VERSION; }

extern "C" const char* emo_last_error_string(void) { return g_err; }

extern "C" size_t emo_workspace_bytes(int op, int precision, int B, int T, int U1, int J, int V) {
    if (B <= 0 || T <= 0) return 0;
    switch (op) {
        case EMO_OP_RNNT_JOINT_FWD:
        case EMO_OP_RNNT_JOINT_BWD:
            if (U1 <= 0 || J <= 0 || V <= 0) return 0;
            return precision == EMO_PREC_BF16 ? joint_bf16_workspace(op, B, T, U1, J, V)
                                              : joint_f32_workspace(op, B, T, U1, J, V);
        default:
            return 0;  // CTC takes its scratch as explicit arguments
    }
}

extern "C" int emo_rnnt_joint_supported(int precision, int B, int T, int U1, int J, int V) {
    if (B <= 0 || T <= 0 || U1 <= 0 || J <= 0 || V <= 0) return 0;
    if (precision == EMO_PREC_FP32) return 1;
    if (precision != EMO_PREC_BF16) return 0;
    return joint_ring_supported(B, T, U1, J, V) ? 1 : 0;
}

extern "C" int emo_launch_count(int op, int precision, int B, int T, int U1, int J, int V) {
    if (B <= 0 || T <= 0) return 0;
    switch (op) {
        case EMO_OP_RNNT_JOINT_FWD:
            return (precision == EMO_PREC_BF16 ? joint_bf16_launches(op, B, T, U1, J, V)
                                               : joint_f32_launches(op, B, T, U1, J, V)) + 2;
        case EMO_OP_RNNT_JOINT_BWD:
            return precision == EMO_PREC_BF16 ? joint_bf16_launches(op, B, T, U1, J, V)
                                              : joint_f32_launches(op, B, T, U1, J, V);
        case EMO_OP_CTC:
            return 3;  // row lse + emission gather, alpha || beta lattices, gradient
        case EMO_OP_RNNT_JOINT_FULL:   // fwd: casts, both projections, joint forward, lattice, posteriors; bwd: ring prep,
                                       // ring kernel, axis reductions (+ d_enc copy / bias sums), d_dec copy + bias sums,
                                       // both d_x GEMMs, both d_W GEMMs
            return 11 + ((V + 31) / 32 * 32 != V ? 1 : 0);
        case EMO_OP_CTC_HEAD:   // J = He.  fwd: prep, 2 casts, label-row gather, joint forward, emissions, lattices; bwd: prep,
                                // 2 casts, gather, per-frame scale, ring prep, ring kernel, occupancies, d_eouts, d_W
                                // (+ the vocabulary padding kernel in each)
            return 17 + ((V + 31) / 32 * 32 != V ? 2 : 0);
        default:
            return 0;
    }
}

extern "C" int emo_rnnt_joint_fwd(const float* enc_proj, const float* dec_proj, const float* w_out,
                                  const float* b_out, const int* labels, const int* tlen,
                                  const int* ulen, int B, int T, int U1, int J, int V, int blank,
                                  int precision, float* lp2, float* lse, void* ws, size_t ws_bytes,
                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == EMO_PREC_FP32)
        return joint_fwd_f32(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, B, T, U1, J, V,
                             blank, lp2, lse, ws, ws_bytes, st);
    if (precision == EMO_PREC_BF16)
        return joint_fwd_bf16(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, B, T, U1, J, V,
                              blank, lp2, lse, ws, ws_bytes, st);
    set_error("joint_fwd: unknown precision %d", precision);
    return EMO_BAD_ARG;
}

extern "C" int emo_rnnt_joint_bwd(const float* enc_proj, const float* dec_proj, const float* w_out,
                                  const float* b_out, const int* labels, const int* tlen,
                                  const int* ulen, const float* lse, const float* lp2, const float* gamma2,
                                  const float* grad_cost, const float* grad_lse,
                                  int B, int T, int U1, int J, int V, int blank, int precision, float* d_enc_proj, float* d_dec_proj,
                                  float* d_w_out, float* d_b_out, void* ws, size_t ws_bytes,
                                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == EMO_PREC_FP32)
        return joint_bwd_f32(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, lse, gamma2,
                             grad_cost, grad_lse, B, T, U1, J, V, blank, d_enc_proj, d_dec_proj, d_w_out,
                             d_b_out, ws, ws_bytes, st);
    if (precision == EMO_PREC_BF16)
        return joint_bwd_bf16(enc_proj, dec_proj, w_out, b_out, labels, tlen, ulen, lse, lp2, gamma2,
                              grad_cost, grad_lse, B, T, U1, J, V, blank, d_enc_proj, d_dec_proj, d_w_out,
                              d_b_out, ws, ws_bytes, st);
    set_error("joint_bwd: unknown precision %d", precision);
    return EMO_BAD_ARG;
}

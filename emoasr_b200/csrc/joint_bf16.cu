// placeholder until the tcgen05 path lands (replaced in the next milestone)
#include "common.cuh"
namespace emo {
size_t joint_bf16_workspace(int, int, int, int, int, int) { return 256; }
int joint_fwd_bf16(const float*, const float*, const float*, const float*, const int*, const int*,
                   const int*, int, int, int, int, int, int, float*, float*, void*, size_t,
                   cudaStream_t) {
    set_error("joint_fwd(bf16): not built");
    return EMO_UNSUPPORTED_SHAPE;
}
int joint_bwd_bf16(const float*, const float*, const float*, const float*, const int*, const int*,
                   const int*, const float*, const float*, const float*, int, int, int, int, int,
                   int, float*, float*, float*, float*, void*, size_t, cudaStream_t) {
    set_error("joint_bwd(bf16): not built");
    return EMO_UNSUPPORTED_SHAPE;
}
}  // namespace emo

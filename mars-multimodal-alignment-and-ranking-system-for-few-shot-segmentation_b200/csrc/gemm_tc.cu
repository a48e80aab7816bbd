// placeholder until the tcgen05 kernel lands
#include "gemm_common.cuh"
namespace marsb200 {
int gemm_tcgen05(const float*, const float*, const float*, const float*, int, int64_t, int64_t, int64_t,
                 const GemmEpilogue&, cudaStream_t) {
    return fail(MARSB200_ERR_UNSUPPORTED, "%s: tcgen05 contraction not built", "gemm_tcgen05");
}
int pairwise_mma(const uint32_t*, int, int, int64_t, int32_t*, cudaStream_t) {
    return fail(MARSB200_ERR_UNSUPPORTED, "%s: tcgen05 pairwise kernel not built", "pairwise_mma");
}
}  // namespace marsb200

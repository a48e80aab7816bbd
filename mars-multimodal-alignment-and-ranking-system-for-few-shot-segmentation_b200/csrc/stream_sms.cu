// SM count a launch on `stream` can use.  A stream created inside a CUDA green context (a spatial partition of the
// device: cuDevSmResourceSplitByCount + cuGreenCtxCreate + cuGreenCtxStreamCreate) only sees that partition's SMs;
// the persistent kernels size their grids from it so that the HBM-bound ingest can run on one partition while the
// tensor-core kernels run on the other (episodes.py).  Plain streams get the device's SM count.
#include <cuda.h>

#include "common.cuh"

namespace marsb200 {

typedef CUresult (*StreamGetGreenCtxFn)(CUstream, CUgreenCtx*);
typedef CUresult (*GreenCtxGetDevResourceFn)(CUgreenCtx, CUdevResource*, CUdevResourceType);

int sms_for_stream(cudaStream_t stream, int* out) {
    int device_sms = 0;  // of the calling thread's current device (cached per device)
    MARS_CUDA_OK(device_sm_count(&device_sms));
    // driver entry points are process-wide; the lookup is idempotent, the flag is published last
    static StreamGetGreenCtxFn get_green = nullptr;
    static GreenCtxGetDevResourceFn get_res = nullptr;
    static std::atomic<bool> looked_up{false};
    if (!looked_up.load(std::memory_order_acquire)) {
        void *p0 = nullptr, *p1 = nullptr;
        cudaDriverEntryPointQueryResult q0, q1;
        if (cudaGetDriverEntryPoint("cuStreamGetGreenCtx", &p0, cudaEnableDefault, &q0) == cudaSuccess &&
            q0 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuGreenCtxGetDevResource", &p1, cudaEnableDefault, &q1) == cudaSuccess &&
            q1 == cudaDriverEntryPointSuccess) {
            get_green = reinterpret_cast<StreamGetGreenCtxFn>(p0);
            get_res = reinterpret_cast<GreenCtxGetDevResourceFn>(p1);
        }
        (void)cudaGetLastError();
        looked_up.store(true, std::memory_order_release);
    }
    *out = device_sms;
    if (!get_green || !stream) return MARSB200_OK;
    // legacy / per-thread default stream handles are not real streams
    if (stream == cudaStreamLegacy || stream == cudaStreamPerThread) return MARSB200_OK;
    CUgreenCtx g = nullptr;
    if (get_green(reinterpret_cast<CUstream>(stream), &g) != CUDA_SUCCESS || !g) return MARSB200_OK;
    CUdevResource res;
    memset(&res, 0, sizeof(res));
    if (get_res(g, &res, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS && res.sm.smCount > 0 &&
        (int)res.sm.smCount <= device_sms)
        *out = (int)res.sm.smCount;
    return MARSB200_OK;
}

}  // namespace marsb200

extern "C" int marsb200_stream_sm_count(void* stream, int* count_host) {
    MARS_REQUIRE(count_host, "null pointer");
    return marsb200::sms_for_stream(marsb200::as_stream(stream), count_host);
}

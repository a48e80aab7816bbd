// SM count a launch on `stream` can use.  A stream created inside a CUDA green context (a spatial partition of the
// device: cuDevSmResourceSplitByCount + cuGreenCtxCreate + cuGreenCtxStreamCreate) only sees that partition's SMs;
// the persistent kernels size their grids from it so that the HBM-bound ingest can run on one partition while the
// tensor-core kernels run on the other (episodes.py).  Plain streams get the device's SM count.
#include <cuda.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace marsb200 {

// Optional per-stream caps (marsb200_stream_set_sm_cap): a caller that wants a persistent kernel to leave SMs to the
// kernels of other streams WITHOUT a spatial partition - the single-episode schedule caps the alignment streams so the
// contractions do not take the whole device away from the mask ingest running beside them.
static std::mutex g_cap_mutex;
static std::vector<std::pair<cudaStream_t, int>> g_caps;
static std::atomic<int> g_cap_count{0};

static int cap_for_stream(cudaStream_t stream) {
    if (g_cap_count.load(std::memory_order_acquire) == 0) return 0;
    std::lock_guard<std::mutex> lock(g_cap_mutex);
    for (const auto& c : g_caps)
        if (c.first == stream) return c.second;
    return 0;
}

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
    if (const int cap = cap_for_stream(stream)) *out = cap < device_sms ? cap : device_sms;
    if (!get_green || !stream) return MARSB200_OK;
    // legacy / per-thread default stream handles are not real streams
    if (stream == cudaStreamLegacy || stream == cudaStreamPerThread) return MARSB200_OK;
    CUgreenCtx g = nullptr;
    if (get_green(reinterpret_cast<CUstream>(stream), &g) != CUDA_SUCCESS || !g) return MARSB200_OK;
    CUdevResource res;
    memset(&res, 0, sizeof(res));
    if (get_res(g, &res, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS && res.sm.smCount > 0 &&
        (int)res.sm.smCount <= *out)
        *out = (int)res.sm.smCount;
    return MARSB200_OK;
}

}  // namespace marsb200

extern "C" int marsb200_stream_sm_count(void* stream, int* count_host) {
    MARS_REQUIRE(count_host, "null pointer");
    return marsb200::sms_for_stream(marsb200::as_stream(stream), count_host);
}

extern "C" int marsb200_stream_set_sm_cap(void* stream, int cap) {
    MARS_REQUIRE(cap >= 0, "cap must be >= 0 (0 removes it)");
    cudaStream_t s = marsb200::as_stream(stream);
    std::lock_guard<std::mutex> lock(marsb200::g_cap_mutex);
    auto& caps = marsb200::g_caps;
    for (size_t i = 0; i < caps.size(); ++i)
        if (caps[i].first == s) {
            if (cap) caps[i].second = cap;
            else caps.erase(caps.begin() + i);
            marsb200::g_cap_count.store((int)caps.size(), std::memory_order_release);
            return MARSB200_OK;
        }
    if (cap) caps.emplace_back(s, cap);
    marsb200::g_cap_count.store((int)caps.size(), std::memory_order_release);
    return MARSB200_OK;
}

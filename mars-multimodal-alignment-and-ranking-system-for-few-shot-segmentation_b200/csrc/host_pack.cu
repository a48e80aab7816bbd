// Host side of the proposal ingest: float32 / uint8 masks in HOST memory -> the packed bit layout of marsb200_pack_masks,
// by a team of host threads, so that 1/32 of the bytes cross PCIe (the reference's proposals are CPU tensors,
// main_MARS.py:62, FilteringMergingModule.py:73).  A B200's PCIe 5 link carries 55 GB/s of float32 masks = 50 c2 episodes/s;
// sixteen host cores read them at 116 GB/s.  This is a format conversion in front of the copy, not a scoring path: every
// score, rank and merge still comes from the device kernels.
#include <immintrin.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace marsb200 {

// 32 pixels -> one word, bit k = pixel 32 w + k is set (> 0; NaN counts as clear, like the device kernel's comparison)
__attribute__((target("avx2"))) static void host_pack_f32_avx2(const float* src, uint32_t* dst, int64_t full_words) {
    const __m256 zero = _mm256_setzero_ps();
    for (int64_t w = 0; w < full_words; ++w) {
        const float* p = src + w * 32;
        const uint32_t m0 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p), zero, _CMP_GT_OQ));
        const uint32_t m1 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 8), zero, _CMP_GT_OQ));
        const uint32_t m2 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 16), zero, _CMP_GT_OQ));
        const uint32_t m3 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 24), zero, _CMP_GT_OQ));
        dst[w] = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
    }
}

__attribute__((target("avx2"))) static void host_pack_u8_avx2(const uint8_t* src, uint32_t* dst, int64_t full_words) {
    const __m256i zero = _mm256_setzero_si256();
    for (int64_t w = 0; w < full_words; ++w) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + w * 32));
        dst[w] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero));
    }
}

template <typename T>
static void host_pack_scalar(const T* src, uint32_t* dst, int64_t px0, int64_t px1) {  // pixels [px0, px1) of one mask, px0 % 32 == 0
    for (int64_t w = px0 / 32; w * 32 < px1; ++w) {
        uint32_t word = 0;
        const int64_t lim = std::min<int64_t>(32, px1 - w * 32);
        for (int64_t k = 0; k < lim; ++k) word |= (src[w * 32 + k] > (T)0 ? 1u : 0u) << k;
        dst[w] = word;
    }
}

template <typename T>
static void host_pack_rows(const T* masks, int64_t m0, int64_t m1, int64_t HW, int64_t wpm, uint32_t* bits, bool avx2) {
    const int64_t full = HW / 32;
    for (int64_t m = m0; m < m1; ++m) {
        const T* src = masks + m * HW;
        uint32_t* dst = bits + m * wpm;
        int64_t done = 0;
        if (avx2) {
            if (sizeof(T) == 4) host_pack_f32_avx2(reinterpret_cast<const float*>(src), dst, full);
            else host_pack_u8_avx2(reinterpret_cast<const uint8_t*>(src), dst, full);
            done = full;
        }
        host_pack_scalar(src, dst, done * 32, HW);
        const int64_t used = (HW + 31) / 32;
        if (used < wpm) std::memset(dst + used, 0, (size_t)(wpm - used) * 4);  // zero tail of the 128-byte aligned row
    }
}

// A team of host threads that outlives the call (an ingest step packs a few hundred masks in ~10 ms: spawning sixteen
// threads per call would be a visible share of that).  Workers sleep on a condition variable between calls; the pool is
// never destroyed (detached threads), calls are serialised.
class HostTeam {
public:
    void run(int threads, const std::function<void(int)>& job) {
        std::lock_guard<std::mutex> call(call_);
        {
            std::unique_lock<std::mutex> lk(m_);
            if (owner_ != getpid()) {  // a forked child inherits the counters but none of the worker threads
                owner_ = getpid();
                spawned_ = 0;
            }
            while ((int)spawned_ < threads - 1) {
                const int index = ++spawned_;
                std::thread([this, index] { worker(index); }).detach();
            }
            job_ = &job;
            active_ = threads;
            pending_ = threads - 1;
            ++generation_;
        }
        cv_work_.notify_all();
        job(0);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void worker(int index) {
        unsigned seen = 0;
        while (true) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (index < active_) job = job_;
            }
            if (job) {
                (*job)(index);
                std::unique_lock<std::mutex> lk(m_);
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }
    std::mutex call_, m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    unsigned generation_ = 0;
    int spawned_ = 0, active_ = 0, pending_ = 0;
    pid_t owner_ = 0;
};

static HostTeam& host_team() {
    static HostTeam* team = new HostTeam();  // leaked on purpose: its detached workers may outlive static destructors
    return *team;
}

}  // namespace marsb200

using namespace marsb200;

extern "C" int marsb200_host_pack_masks(const void* masks, int mask_dtype, int64_t n, int64_t HW, uint32_t* bits, int threads) {
    MARS_REQUIRE(masks && bits, "null pointer");
    MARS_REQUIRE(n > 0 && HW > 0, "empty input");
    MARS_REQUIRE(mask_dtype == MARSB200_MASK_F32 || mask_dtype == MARSB200_MASK_U8, "mask_dtype");
    const int64_t wpm = marsb200_words_per_mask(HW);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<int64_t>(std::min(threads, 256), n);
    const bool avx2 = __builtin_cpu_supports("avx2");
    const std::function<void(int)> work = [&](int t) {
        const int64_t m0 = n * t / threads, m1 = n * (t + 1) / threads;
        if (mask_dtype == MARSB200_MASK_F32) host_pack_rows(static_cast<const float*>(masks), m0, m1, HW, wpm, bits, avx2);
        else host_pack_rows(static_cast<const uint8_t*>(masks), m0, m1, HW, wpm, bits, avx2);
    };
    if (threads == 1) work(0);
    else host_team().run(threads, work);
    return MARSB200_OK;
}

// host-side f32 -> bit packing rate with N threads (is it faster than pushing the floats over PCIe at 55 GB/s?)
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void pack_range(const float* src, uint32_t* dst, size_t words) {
    const __m256 zero = _mm256_setzero_ps();
    for (size_t w = 0; w < words; ++w) {
        const float* p = src + w * 32;
        uint32_t m0 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p), zero, _CMP_GT_OQ));
        uint32_t m1 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 8), zero, _CMP_GT_OQ));
        uint32_t m2 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 16), zero, _CMP_GT_OQ));
        uint32_t m3 = (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + 24), zero, _CMP_GT_OQ));
        dst[w] = m0 | (m1 << 8) | (m2 << 16) | (m3 << 24);
    }
}
int main(int argc, char** argv) {
    const size_t bytes = (size_t)2 << 30;  // 2 GiB of floats = two c2 episodes
    const size_t n = bytes / 4, words = n / 32;
    float* src = (float*)aligned_alloc(4096, bytes);
    uint32_t* dst = (uint32_t*)aligned_alloc(4096, words * 4);
    printf("hardware threads: %u\n", std::thread::hardware_concurrency());
    { std::vector<std::thread> th; int T = 16; for (int t = 0; t < T; ++t) th.emplace_back([=] { size_t a = n * t / T, b = n * (t + 1) / T; for (size_t i = a; i < b; ++i) src[i] = (i * 2654435761u >> 13) & 1 ? 1.f : 0.f; }); for (auto& x : th) x.join(); }
    for (int T : {1, 4, 8, 16, 32, 64}) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([=] { size_t a = words * t / T, b = words * (t + 1) / T; pack_range(src + a * 32, dst + a, b - a); });
            for (auto& x : th) x.join();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (s < best) best = s;
        }
        printf("threads %2d: %.1f ms for 2 GiB -> %.1f GB/s of float32 read (%.0f c2 episodes/s)\n", T, best * 1e3, bytes / best / 1e9, 2.0 / best * (2147483648.0 / 2 / 1073741824.0));
    }
    return 0;
}

// Row / column arg-max and row top-k of the similarity matrix with warp-shuffle reductions:
// the mutual-nearest-neighbour candidates of bidirectional patch matching (BASELINE north-star kernel 1;
// the reference's Matcher uses two exact assignments, matcher/Matcher.py:443-477 - these
// arg-max outputs are the GPU-side candidates / seeds, not a replacement for the assignment).
#include "common.cuh"

namespace marsb200 {

constexpr int MATCH_MAX_K = 8;

// one warp per row: top-k (k <= 8) by repeated warp arg-max; ties -> lowest column index (torch.topk on CPU
// returns the first occurrence for distinct values; equal values are unordered there)
__global__ void __launch_bounds__(256) row_topk_kernel(const float* __restrict__ S, int64_t total_rows, int N, int k,
                                                       float* __restrict__ vals, int32_t* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= total_rows) return;
    const float* s = S + row * N;
    int taken[MATCH_MAX_K];
    for (int r = 0; r < k; ++r) {
        float best = -INFINITY;
        int best_j = 0x7fffffff;
        for (int j = lane; j < N; j += 32) {
            bool skip = false;
            for (int q = 0; q < r; ++q) skip |= (taken[q] == j);
            const float v = s[j];
            if (!skip && (v > best || (v == best && j < best_j))) {
                best = v;
                best_j = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
            if (ov > best || (ov == best && oj < best_j)) {
                best = ov;
                best_j = oj;
            }
        }
        taken[r] = best_j;
        if (lane == 0) {
            vals[row * k + r] = best;
            idx[row * k + r] = best_j;
        }
    }
}

// column arg-max over the rows selected by `row_mask` (all rows when null): a warp owns 32 columns, each lane
// walks its column (coalesced 128-byte row segments), no cross-lane step needed
__global__ void __launch_bounds__(256) col_argmax_kernel(const float* __restrict__ S, const uint8_t* __restrict__ row_mask,
                                                         int M, int N, float* __restrict__ vals,
                                                         int32_t* __restrict__ idx) {
    const int64_t e = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    const float* s = S + e * (int64_t)M * N + col;
    const uint8_t* mk = row_mask ? row_mask + e * M : nullptr;
    float best = -INFINITY;
    int best_i = -1;
    for (int i = 0; i < M; ++i) {
        if (mk && !mk[i]) continue;
        const float v = s[(int64_t)i * N];
        if (v > best) {  // strict: first (lowest) row wins ties
            best = v;
            best_i = i;
        }
    }
    vals[e * N + col] = best;
    idx[e * N + col] = best_i;
}

}  // namespace marsb200

using namespace marsb200;

extern "C" int marsb200_match_argmax(const float* sim, const uint8_t* row_mask, int E, int M, int N, int k,
                                     float* row_vals, int32_t* row_idx, float* col_vals, int32_t* col_idx,
                                     void* stream) {
    MARS_REQUIRE(sim, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && M > 0 && N > 0, "shape");
    MARS_REQUIRE((row_vals == nullptr) == (row_idx == nullptr) && (col_vals == nullptr) == (col_idx == nullptr),
                 "value / index outputs go together");
    MARS_REQUIRE(row_vals || col_vals, "no output requested");
    if (row_vals) {
        MARS_REQUIRE(k >= 1 && k <= MATCH_MAX_K && k <= N, "1 <= k <= min(8, N)");
        const int64_t rows = (int64_t)E * M;
        row_topk_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, as_stream(stream)>>>(sim, rows, N, k, row_vals, row_idx);
        MARS_LAUNCH_OK();
    }
    if (col_vals) {
        col_argmax_kernel<<<dim3(ceil_div(N, 128), E), 128, 0, as_stream(stream)>>>(sim, row_mask, M, N, col_vals, col_idx);
        MARS_LAUNCH_OK();
    }
    return MARSB200_OK;
}

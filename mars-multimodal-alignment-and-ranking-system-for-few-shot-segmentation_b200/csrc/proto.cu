// Mask-pooled features and masked similarity statistics (north-star kernel 2):
//   * prototypes  P[e, p, :] = mean over the patches of pooled mask p of the (unnormalised) patch features
//     (matcher/Matcher.py:1076-1079) as ONE masks-by-features contraction [P, N] x [N, C] on the tensor cores:
//     the 0/1 mask matrix is exact in tf32 (its residual operand is zero), the features go through the same
//     error-compensated split as every other contraction of the path;
//   * mean / max / unbiased std of the similarity sub-matrix S[row mask][:, column mask] (Matcher.py:1072-1085);
//   * the mean over masked support rows of S (get_ref_to_target_similarity, Matcher.py:593-611).
#include "gemm_common.cuh"

namespace marsb200 {

// pooled bitmaps [E, P, npw] -> A [E, pad_rows(P), k_pad] fp32 0/1 (zero padded) and the per-mask counts
__global__ void __launch_bounds__(256) expand_bitmap_kernel(const uint32_t* __restrict__ pooled, int P, int N, int npw,
                                                            int64_t p_pad, int64_t k_pad, float* __restrict__ a,
                                                            int32_t* __restrict__ count) {
    const int64_t e = blockIdx.y, r = blockIdx.x;  // r over padded rows
    float* dst = a + (e * p_pad + r) * k_pad;
    if (r >= P) {
        for (int64_t c = threadIdx.x; c < k_pad; c += blockDim.x) dst[c] = 0.f;
        return;
    }
    const uint32_t* src = pooled + (e * P + r) * npw;
    int cnt = 0;
    for (int64_t c = threadIdx.x; c < k_pad; c += blockDim.x) {
        const bool on = c < N && ((src[c >> 5] >> (c & 31)) & 1u);
        dst[c] = on ? 1.f : 0.f;
        cnt += on ? 1 : 0;
    }
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) count[e * P + r] = s_cnt;
}

// feats [E, N, C] -> bt, bt_lo [E, pad_rows(C), k_pad]: transposed (k = patch index contiguous) + tf32 residual
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ feats, int N, int C, int64_t c_pad,
                                                              int64_t k_pad, float* __restrict__ bt,
                                                              float* __restrict__ bt_lo) {
    __shared__ float tile[32][33];
    const int64_t e = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, c = c0 + tx;
        tile[i][tx] = (n < N && c < C) ? feats[(e * N + n) * (int64_t)C + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, n = n0 + tx;
        if (c < c_pad && n < k_pad) {
            const float v = tile[tx][i];
            bt[(e * c_pad + c) * k_pad + n] = v;
            bt_lo[(e * c_pad + c) * k_pad + n] = tf32_residual(v);
        }
    }
}

__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ x, const int32_t* __restrict__ count,
                                                         int64_t rows, int64_t C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const int cnt = count[i / C];
    x[i] = cnt > 0 ? __fdiv_rn(x[i], (float)cnt) : NAN;  // mean of an empty selection is NaN in torch
}

// one CTA per episode: mean, max, unbiased std and count of S over the selected rows x columns
__global__ void __launch_bounds__(512) masked_stats_kernel(const float* __restrict__ S, const uint8_t* __restrict__ row_mask,
                                                            const uint8_t* __restrict__ col_mask, int64_t M, int64_t N,
                                                            double* __restrict__ out) {
    __shared__ double s_sum[16], s_sq[16];
    __shared__ float s_max[16];
    __shared__ long long s_cnt[16];
    const int64_t e = blockIdx.x;
    const float* s = S + e * M * N;
    const uint8_t* rm = row_mask + e * M;
    const uint8_t* cm = col_mask + e * N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double sum = 0.0, sq = 0.0;
    float mx = -INFINITY;
    long long cnt = 0;
    for (int64_t r = warp; r < M; r += nw) {
        if (!rm[r]) continue;
        for (int64_t c = lane; c < N; c += 32)
            if (cm[c]) {
                const float v = s[r * N + c];
                sum += (double)v;
                sq += (double)v * (double)v;
                mx = fmaxf(mx, v);
                ++cnt;
            }
    }
    sum = warp_sum(sum);
    sq = warp_sum(sq);
    mx = warp_max(mx);
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        s_sum[warp] = sum;
        s_sq[warp] = sq;
        s_max[warp] = mx;
        s_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        float m = -INFINITY;
        long long n = 0;
        for (int w = 0; w < nw; ++w) {
            a += s_sum[w];
            b += s_sq[w];
            m = fmaxf(m, s_max[w]);
            n += s_cnt[w];
        }
        const double mean = n > 0 ? a / (double)n : nan("");
        out[4 * e + 0] = mean;
        out[4 * e + 1] = n > 0 ? (double)m : 0.0;  // the reference reports 0 for an empty selection (Matcher.py:1083)
        out[4 * e + 2] = n > 1 ? sqrt(fmax(0.0, (b - (double)n * mean * mean) / (double)(n - 1))) : nan("");
        out[4 * e + 3] = (double)n;
    }
}

// out[e, c] = mean over the selected rows of S[e, :, c]; one warp per column block of 32, rows strided over the CTA
__global__ void __launch_bounds__(256) masked_row_mean_kernel(const float* __restrict__ S, const uint8_t* __restrict__ row_mask,
                                                               int64_t M, int64_t N, float* __restrict__ out) {
    __shared__ double s_part[8][32];
    __shared__ int s_cnt[8];
    const int64_t e = blockIdx.y;
    const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5;
    const float* s = S + e * M * N;
    const uint8_t* rm = row_mask + e * M;
    double acc = 0.0;
    int cnt = 0;
    for (int64_t r = warp; r < M; r += 8)
        if (rm[r]) {
            ++cnt;
            if (c < N) acc += (double)s[r * N + c];
        }
    s_part[warp][threadIdx.x & 31] = acc;
    if ((threadIdx.x & 31) == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (warp == 0 && c < N) {
        double tot = 0.0;
        int n = 0;
        for (int w = 0; w < 8; ++w) {
            tot += s_part[w][threadIdx.x];
            n += s_cnt[w];
        }
        out[e * N + c] = n > 0 ? (float)(tot / (double)n) : NAN;
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int64_t marsb200_masked_feature_means_workspace_bytes(int E, int P, int N, int C) {
    if (E <= 0 || P <= 0 || N <= 0 || C <= 0) return 0;
    const int64_t p_pad = marsb200_pad_rows(P), c_pad = marsb200_pad_rows(C), k_pad = marsb200_pad_k(N);
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    return 2 * al((int64_t)E * p_pad * k_pad * 4) + 2 * al((int64_t)E * c_pad * k_pad * 4) + al((int64_t)E * P * 4);
}

int marsb200_masked_feature_means(const uint32_t* pooled, const float* feats, int E, int P, int N, int C, float* out,
                                  void* workspace, int64_t workspace_bytes, int backend, void* stream) {
    MARS_REQUIRE(pooled && feats && out && workspace, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && P > 0 && N > 0 && C > 0, "shape");
    MARS_REQUIRE(workspace_bytes >= marsb200_masked_feature_means_workspace_bytes(E, P, N, C), "workspace too small");
    MARS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int64_t p_pad = marsb200_pad_rows(P), c_pad = marsb200_pad_rows(C), k_pad = marsb200_pad_k(N);
    auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
    char* w = reinterpret_cast<char*>(workspace);
    float* a = reinterpret_cast<float*>(w);
    float* a_lo = reinterpret_cast<float*>(w + al((int64_t)E * p_pad * k_pad * 4));
    float* bt = reinterpret_cast<float*>(w + 2 * al((int64_t)E * p_pad * k_pad * 4));
    float* bt_lo = reinterpret_cast<float*>(w + 2 * al((int64_t)E * p_pad * k_pad * 4) + al((int64_t)E * c_pad * k_pad * 4));
    int32_t* count = reinterpret_cast<int32_t*>(w + 2 * al((int64_t)E * p_pad * k_pad * 4) + 2 * al((int64_t)E * c_pad * k_pad * 4));
    const int npw = ceil_div(N, 32);
    MARS_CUDA_OK(cudaMemsetAsync(a_lo, 0, (size_t)E * p_pad * k_pad * 4, s));  // 0/1 is exact in tf32
    expand_bitmap_kernel<<<dim3((unsigned)p_pad, E), 256, 0, s>>>(pooled, P, N, npw, p_pad, k_pad, a, count);
    MARS_LAUNCH_OK();
    transpose_split_kernel<<<dim3((unsigned)(k_pad / 32), (unsigned)(c_pad / 32), E), 256, 0, s>>>(feats, N, C, c_pad, k_pad,
                                                                                                  bt, bt_lo);
    MARS_LAUNCH_OK();
    GemmEpilogue ep{};
    ep.out0 = out;
    ep.M = P;
    ep.N = C;
    ep.ld_out = C;
    ep.tiles_m = (int)(p_pad / GEMM_BM);
    const GemmOperand oa{a, a_lo, p_pad, k_pad, p_pad * k_pad}, ob{bt, bt_lo, c_pad, k_pad, c_pad * k_pad};
    int rc;
    if (backend == MARSB200_GEMM_SIMT)
        rc = gemm_simt(oa, ob, E, P, C, N, ep, s);
    else if (backend == MARSB200_GEMM_TCGEN05)
        rc = gemm_tcgen05(oa, ob, E, P, C, N, ep, s);
    else
        return fail(MARSB200_ERR_ARG, "%s: unknown backend %lld", "marsb200_masked_feature_means", backend);
    if (rc != MARSB200_OK) return rc;
    scale_rows_kernel<<<(unsigned)ceil_div64((int64_t)E * P * C, 256), 256, 0, s>>>(out, count, (int64_t)E * P, C);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_masked_sim_stats(const float* sim, const uint8_t* row_mask, const uint8_t* col_mask, int E, int64_t M,
                              int64_t N, double* out, void* stream) {
    MARS_REQUIRE(sim && row_mask && col_mask && out, "null pointer");
    MARS_REQUIRE(E > 0 && M > 0 && N > 0, "shape");
    masked_stats_kernel<<<E, 512, 0, as_stream(stream)>>>(sim, row_mask, col_mask, M, N, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_masked_row_mean(const float* sim, const uint8_t* row_mask, int E, int64_t M, int64_t N, float* out,
                             void* stream) {
    MARS_REQUIRE(sim && row_mask && out, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && M > 0 && N > 0, "shape");
    masked_row_mean_kernel<<<dim3((unsigned)ceil_div64(N, 32), E), 256, 0, as_stream(stream)>>>(sim, row_mask, M, N, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

// Prior-information refinement (A4): attention mean, box mask by 8-connected components,
// Sinkhorn-1 normalisation, R = max(D, D D^T) through the shared contraction, and the refinement
// applied as two matrix-vector products, R (R (B * prior)) == ((R R) * B) prior.
#include <cuda_fp16.h>

#include "gemm_common.cuh"

namespace marsb200 {

int minmax_rows(float* v, int E, int64_t n, cudaStream_t s);  // vva.cu

// --------------------------------------------------------------------------------------------
// attention mean over (layers, heads); HBM-bound: every map element is read exactly once.
// --------------------------------------------------------------------------------------------
constexpr int MAX_ATTN_MAPS = 64;
struct AttnPtrs {
    const void* p[MAX_ATTN_MAPS];
};

template <typename T>
__global__ void __launch_bounds__(256) attn_mean_kernel(AttnPtrs maps, int n_maps, int heads, int T_tokens, int skip,
                                                        float* __restrict__ out, int64_t ld_out) {
    const int N = T_tokens - skip;
    const int row = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    float acc = 0.f;
    for (int l = 0; l < n_maps; ++l) {
        const T* base = reinterpret_cast<const T*>(maps.p[l]);
        for (int h = 0; h < heads; ++h)
            acc += (float)base[((int64_t)h * T_tokens + (row + skip)) * T_tokens + (col + skip)];
    }
    float mean = acc / (float)(n_maps * heads);
    // torch.mean of fp16 maps returns fp16 (fp32 accumulation, rounded once) before the .float()
    if (sizeof(T) == 2) mean = __half2float(__float2half_rn(mean));
    out[(int64_t)row * ld_out + col] = mean;
}

// fp16 maps with an even token count: every row starts 4-byte aligned, so a thread reads an aligned __half2 (two
// source columns) per map and head - 128 bytes per warp and load like the fp32 kernel instead of 64.  Same summation
// order per element as attn_mean_kernel<__half>: identical results.
__global__ void __launch_bounds__(256) attn_mean_h2_kernel(AttnPtrs maps, int n_maps, int heads, int T_tokens, int skip,
                                                           float* __restrict__ out, int64_t ld_out) {
    const int row = blockIdx.y;
    const int j0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + (skip & ~1);  // even source column
    if (j0 >= T_tokens) return;
    float a0 = 0.f, a1 = 0.f;
    for (int l = 0; l < n_maps; ++l) {
        const __half* base = reinterpret_cast<const __half*>(maps.p[l]);
#pragma unroll 4
        for (int h = 0; h < heads; ++h) {
            const __half2 v = *reinterpret_cast<const __half2*>(base + ((int64_t)h * T_tokens + (row + skip)) * T_tokens + j0);
            a0 += __low2float(v);
            a1 += __high2float(v);
        }
    }
    const float inv = (float)(n_maps * heads);
    float* dst = out + (int64_t)row * ld_out;
    if (j0 >= skip) dst[j0 - skip] = __half2float(__float2half_rn(a0 / inv));
    dst[j0 + 1 - skip] = __half2float(__float2half_rn(a1 / inv));
}

// --------------------------------------------------------------------------------------------
// box mask: one block per episode.  img = uint8(prior*255); thr = int(threshold * max(img));
// fg = img > thr; 8-connected components by min-label propagation; per component the box
// [x0, min(x0+w, W-1)) x [y0, min(y0+h, H-1)) is filled (PriorInformationRefinementModule.py:56-63,
// 91-122).  Writes B (uint8, optional) and v = B * prior.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) box_mask_kernel(const float* __restrict__ prior, int g, double threshold,
                                                       uint8_t* __restrict__ box_out, float* __restrict__ v_out,
                                                       int32_t* __restrict__ boxes_out, int32_t* __restrict__ count_out) {
    extern __shared__ int s_mem[];
    const int n = g * g;
    int* label = s_mem;          // n
    int* bx0 = label + n;        // n each
    int* bx1 = bx0 + n;
    int* by0 = bx1 + n;
    int* by1 = by0 + n;
    int* boxm = by1 + n;         // n
    __shared__ int s_max;
    const int64_t e = blockIdx.x;
    const float* p = prior + e * n;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    int local_max = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float q = p[i] * 255.0f;
        const int v = (int)fminf(fmaxf(q, 0.f), 255.f);  // astype(uint8) truncation for in-range values
        label[i] = v;                                     // temporarily holds the quantised value
        local_max = max(local_max, v);
    }
    atomicMax(&s_max, local_max);
    __syncthreads();
    const int thr = (int)(threshold * (double)s_max);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        label[i] = (label[i] > thr) ? i : -1;
        bx0[i] = g;
        by0[i] = g;
        bx1[i] = -1;
        by1[i] = -1;
        boxm[i] = 0;
    }
    __syncthreads();
    // min-label propagation over the 8-neighbourhood with pointer jumping
    while (true) {
        int changed = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            int l = label[i];
            if (l < 0) continue;
            const int y = i / g, x = i % g;
            int best = l;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= g || xx < 0 || xx >= g) continue;
                    const int ln = label[yy * g + xx];
                    if (ln >= 0 && ln < best) best = ln;
                }
            const int root = label[best];  // one jump (labels only decrease, always >= 0 for fg)
            if (root >= 0 && root < best) best = root;
            if (best < l) {
                label[i] = best;
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = label[i];
        if (l < 0) continue;
        const int y = i / g, x = i % g;
        atomicMin(&bx0[l], x);
        atomicMax(&bx1[l], x);
        atomicMin(&by0[l], y);
        atomicMax(&by1[l], y);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (label[i] != i) continue;  // roots only
        const int x0 = bx0[i], y0 = by0[i];
        const int x1 = min(bx1[i] + 1, g - 1), y1 = min(by1[i] + 1, g - 1);  // x0 + w clipped, exclusive
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) boxm[y * g + x] = 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (box_out) box_out[e * n + i] = (uint8_t)boxm[i];
        if (v_out) v_out[e * n + i] = boxm[i] ? p[i] : 0.f;
    }
    // optional box list (the reference's `_scoremap2bbox` return value, :112-122): one (x0, y0, x1, y1) per component,
    // clipped like the fill above, in raster order of each component's first pixel.  Diagnostic output, off the hot path.
    if (boxes_out && threadIdx.x == 0) {
        int k = 0;
        for (int i = 0; i < n; ++i) {
            if (label[i] != i) continue;
            int32_t* b = boxes_out + ((int64_t)e * n + k) * 4;
            b[0] = bx0[i];
            b[1] = by0[i];
            b[2] = min(bx1[i] + 1, g - 1);
            b[3] = min(by1[i] + 1, g - 1);
            ++k;
        }
        count_out[e] = k;
    }
}

// --------------------------------------------------------------------------------------------
// Sinkhorn-1: column sums in two deterministic stages, then one block per row.
// --------------------------------------------------------------------------------------------
constexpr int COLSUM_SPLITS = 16;

__global__ void __launch_bounds__(128) colsum_partial_kernel(const float* __restrict__ attn, int64_t ld, int N,
                                                             double* __restrict__ partial) {
    const int64_t e = blockIdx.z;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int split = blockIdx.y;
    if (col >= N) return;
    const int rows_per = ceil_div(N, COLSUM_SPLITS);
    const int r0 = split * rows_per, r1 = min(r0 + rows_per, N);
    const float* a = attn + e * N * ld;
    double acc = 0.0;
    for (int r = r0; r < r1; ++r) acc += (double)a[(int64_t)r * ld + col];
    partial[(e * COLSUM_SPLITS + split) * N + col] = acc;
}

__global__ void colsum_final_kernel(const double* __restrict__ partial, int N, float* __restrict__ colsum) {
    const int64_t e = blockIdx.y;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= N) return;
    double cs = 0.0;
#pragma unroll
    for (int s = 0; s < COLSUM_SPLITS; ++s) cs += partial[(e * COLSUM_SPLITS + s) * N + col];
    // stored as the correctly rounded reciprocal: the row kernel multiplies (its two IEEE divisions per element made it
    // instruction-issue bound: 70 % issue-active, 3.5 TB/s); a * rcp(c) differs from a / c by less than two ulp
    colsum[e * N + col] = __frcp_rn((float)cs);
}

__global__ void __launch_bounds__(256) row_normalize_kernel(const float* __restrict__ attn, int64_t ld, int N,
                                                            const float* __restrict__ colsum, int64_t k_pad,
                                                            float* __restrict__ D, float* __restrict__ D_lo) {
    __shared__ double s_red[8];
    const int64_t e = blockIdx.y;
    const int row = blockIdx.x;
    const float* a = attn + (e * N + row) * ld;
    float* d = D + (e * N + row) * k_pad;
    float* dl = D_lo + (e * N + row) * k_pad;
    // the row stays in registers between the two passes (N <= 8 * 256 on this path; longer rows go through D)
    constexpr int RN_MAX = 8;
    const bool in_regs = N <= RN_MAX * 256;
    float vals[RN_MAX];
    double acc = 0.0;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < RN_MAX; ++k) {
            const int c = threadIdx.x + k * 256;
            vals[k] = c < N ? a[c] * colsum[e * N + c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < RN_MAX; ++k) acc += (double)vals[k];  // same order as the strided loop below
    } else {
        for (int c = threadIdx.x; c < N; c += blockDim.x) {
            const float v = a[c] * colsum[e * N + c];
            d[c] = v;
            acc += (double)v;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    double total = 0.0;
    for (int w = 0; w < 8; ++w) total += s_red[w];
    const float rs = __frcp_rn((float)total);  // reciprocal of the row sum
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < RN_MAX; ++k) {
            const int64_t c = threadIdx.x + k * 256;
            if (c < k_pad) {
                const float v = c < N ? vals[k] * rs : 0.f;
                d[c] = v;
                dl[c] = tf32_residual(v);
            }
        }
        return;
    }
    for (int64_t c = threadIdx.x; c < k_pad; c += blockDim.x) {
        float v = 0.f;
        if (c < N) v = d[c] * rs;
        d[c] = v;
        dl[c] = tf32_residual(v);
    }
}

// y[e, i] = sum_j max(D[e, i, j], G[e, i, j]) x[e, j] = (R x)_i with R = max(D, D D^T) formed on the fly
// (PriorInformationRefinementModule.py:74-75); one warp per row, double accumulation.  Keeping the max out of
// the contraction's epilogue keeps that epilogue store-only: its loads would queue behind the operand feed.
__global__ void __launch_bounds__(256) matvec_kernel(const float* __restrict__ G, int64_t ld, const float* __restrict__ D,
                                                     int64_t ld_d, int N, const float* __restrict__ x,
                                                     float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t e = blockIdx.y;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const float* r = G + (e * N + row) * ld;
    const float* d = D + (e * N + row) * ld_d;
    const float* xv = x + e * N;
    double acc = 0.0;
    for (int j = lane; j < N; j += 32) acc += (double)fmaxf(d[j], r[j]) * (double)xv[j];
    acc = warp_sum(acc);
    if (lane == 0) y[e * N + row] = (float)acc;
}

struct PirWorkspace {
    double* partial;  // [E, 16, N]
    float* D;         // [E, N, k_pad]: the doubly normalised attention, also the contraction operand
    float* D_lo;      // [E, N, k_pad]: its tf32 residual
    float* R;         // [E, N, N]: G = D D^T (the max with D is taken in the mat-vecs)
    float* v;         // [E, N]
    float* t;         // [E, N]
    float* colsum;    // [E, N]
    int64_t bytes;
};

static PirWorkspace carve(void* base, int E, int64_t N) {
    const int64_t k_pad = marsb200_pad_k(N);
    auto align = [](int64_t b) { return (b + 255) / 256 * 256; };
    char* p = reinterpret_cast<char*>(base);
    int64_t off = 0;
    PirWorkspace w;
    w.partial = reinterpret_cast<double*>(p + off);
    off += align((int64_t)E * COLSUM_SPLITS * N * 8);
    w.D = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * k_pad * 4);
    w.D_lo = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * k_pad * 4);
    w.R = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * N * 4);
    w.v = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * 4);
    w.t = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * 4);
    w.colsum = reinterpret_cast<float*>(p + off);
    off += align((int64_t)E * N * 4);
    w.bytes = off;
    return w;
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int marsb200_attn_mean(const void* const* maps_host, int n_maps, int dtype, int heads, int T, int skip, float* out,
                       int64_t ld_out, void* stream) {
    MARS_REQUIRE(maps_host && out, "null pointer");
    MARS_REQUIRE(n_maps > 0 && n_maps <= MAX_ATTN_MAPS, "1..64 maps");
    MARS_REQUIRE(heads > 0 && T > skip && skip >= 0 && ld_out >= T - skip, "shape");
    MARS_REQUIRE(dtype == 0 || dtype == 1, "dtype (0 fp32, 1 fp16)");
    AttnPtrs ptrs{};
    bool half_aligned = true;
    for (int i = 0; i < n_maps; ++i) {
        MARS_REQUIRE(maps_host[i] != nullptr, "null map");
        ptrs.p[i] = maps_host[i];
        half_aligned = half_aligned && (reinterpret_cast<uintptr_t>(maps_host[i]) & 3) == 0;
    }
    const int N = T - skip;
    dim3 grid(ceil_div(N, 256), N);
    if (dtype == 0)
        attn_mean_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(ptrs, n_maps, heads, T, skip, out, ld_out);
    else if (T % 2 == 0 && half_aligned)
        attn_mean_h2_kernel<<<dim3(ceil_div((T - (skip & ~1)) / 2, 256), N), 256, 0, as_stream(stream)>>>(ptrs, n_maps, heads, T,
                                                                                                         skip, out, ld_out);
    else
        attn_mean_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(ptrs, n_maps, heads, T, skip, out, ld_out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_scoremap_boxes(const float* prior, int E, int g, double box_threshold, int32_t* boxes_out,
                            int32_t* count_out, void* stream) {
    MARS_REQUIRE(prior && boxes_out && count_out, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && g > 0 && g <= 96, "shape (g <= 96)");
    const size_t smem = (size_t)6 * g * g * sizeof(int);
    if (smem > 48 * 1024)
        MARS_CUDA_OK(cudaFuncSetAttribute(box_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    box_mask_kernel<<<E, 256, smem, as_stream(stream)>>>(prior, g, box_threshold, nullptr, nullptr, boxes_out, count_out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int64_t marsb200_pir_workspace_bytes(int E, int64_t N) {
    if (E <= 0 || N <= 0) return 0;
    return carve(nullptr, E, N).bytes;
}

int marsb200_pir_refine(const float* prior, const float* attn, int64_t ld_attn, int E, int g, double box_threshold,
                        int apply_minmax, float* out, uint8_t* box_out, void* workspace, int64_t workspace_bytes,
                        int backend, void* stream) {
    return marsb200_pir_stages(prior, attn, ld_attn, E, g, box_threshold, apply_minmax, out, box_out, workspace,
                               workspace_bytes, backend, MARSB200_PIR_ALL, stream);
}

int marsb200_pir_stages(const float* prior, const float* attn, int64_t ld_attn, int E, int g, double box_threshold,
                        int apply_minmax, float* out, uint8_t* box_out, void* workspace, int64_t workspace_bytes,
                        int backend, int stages, void* stream) {
    MARS_REQUIRE(workspace && (stages & MARSB200_PIR_ALL) != 0, "null workspace / no stage selected");
    MARS_REQUIRE(!(stages & MARSB200_PIR_NORMALISE) || attn, "the normalise stage needs the attention matrix");
    MARS_REQUIRE(!(stages & MARSB200_PIR_APPLY) || (prior && out), "the apply stage needs the prior and the output");
    MARS_REQUIRE(E > 0 && E <= 65535 && g > 0 && g <= 96, "shape (g <= 96)");
    const int N = g * g;
    MARS_REQUIRE(ld_attn >= N, "ld_attn");
    PirWorkspace w = carve(workspace, E, N);
    MARS_REQUIRE(workspace_bytes >= w.bytes, "workspace too small");
    MARS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int64_t n_pad = marsb200_pad_rows(N), k_pad = marsb200_pad_k(N);

    if (stages & MARSB200_PIR_APPLY) {
        const size_t smem = (size_t)6 * N * sizeof(int);
        if (smem > 48 * 1024)
            MARS_CUDA_OK(cudaFuncSetAttribute(box_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        box_mask_kernel<<<E, 256, smem, s>>>(prior, g, box_threshold, box_out, w.v, nullptr, nullptr);
        MARS_LAUNCH_OK();
    }
    if (stages & MARSB200_PIR_NORMALISE) {
        colsum_partial_kernel<<<dim3(ceil_div(N, 128), COLSUM_SPLITS, E), 128, 0, s>>>(attn, ld_attn, N, w.partial);
        MARS_LAUNCH_OK();
        colsum_final_kernel<<<dim3(ceil_div(N, 128), E), 128, 0, s>>>(w.partial, N, w.colsum);
        MARS_LAUNCH_OK();
        row_normalize_kernel<<<dim3((unsigned)N, E), 256, 0, s>>>(attn, ld_attn, N, w.colsum, k_pad, w.D, w.D_lo);
        MARS_LAUNCH_OK();
    }
    if (stages & MARSB200_PIR_CONTRACT) {
    GemmEpilogue ep{};
    ep.out0 = w.R;
    ep.out1 = nullptr;
    ep.maxwith = nullptr;  // R = max(D, G) is applied inside the mat-vecs
    ep.row_fg = nullptr;
    ep.colstats = nullptr;
    ep.M = N;
    ep.N = N;
    ep.ld_out = N;
    ep.ld_max = k_pad;
    ep.tiles_m = (int)(n_pad / GEMM_BM);
    ep.symmetric = (backend == MARSB200_GEMM_TCGEN05) ? 1 : 0;  // D D^T: compute the upper triangle only
    int rc;
    const GemmOperand od{w.D, w.D_lo, N, k_pad, (int64_t)N * k_pad};  // rows beyond N are zero-filled by the loader
    if (backend == MARSB200_GEMM_SIMT)
        rc = gemm_simt(od, od, E, N, N, N, ep, s);
    else if (backend == MARSB200_GEMM_TCGEN05)
        rc = gemm_tcgen05(od, od, E, N, N, N, ep, s);
    else
        return fail(MARSB200_ERR_ARG, "%s: unknown backend %lld", "marsb200_pir_refine", backend);
    if (rc != MARSB200_OK) return rc;
    }
    if (!(stages & MARSB200_PIR_APPLY)) return MARSB200_OK;

    dim3 mv_grid(ceil_div(N, 8), E);
    matvec_kernel<<<mv_grid, 256, 0, s>>>(w.R, N, w.D, k_pad, N, w.v, w.t);
    MARS_LAUNCH_OK();
    matvec_kernel<<<mv_grid, 256, 0, s>>>(w.R, N, w.D, k_pad, N, w.t, out);
    MARS_LAUNCH_OK();
    if (apply_minmax) return minmax_rows(out, E, N, s);
    return MARSB200_OK;
}

}  // extern "C"

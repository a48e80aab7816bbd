// Visual-visual alignment: row normalise (A1), the SIMT validation contraction,
// the similarity entry point (A2/A3), the vva finalisation (A3) and the nearest-resize + min-max
// of the vta map (A5).
#include "gemm_common.cuh"

namespace marsb200 {

// --------------------------------------------------------------------------------------------
// A1: one warp per row: x / max(||x||, 1e-12) into the zero-padded operand layout [E, rows_pad, k_pad], plus the
// tf32 residual of every element (the `lo` operand of the 3xTF32 product).
// Rows >= `rows` and columns >= k of the padded outputs are written as zero.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ x, int64_t ld_x, int64_t rows,
                                                             int64_t k, int64_t rows_pad, int64_t k_pad, int normalize,
                                                             int64_t total_rows, float* __restrict__ out,
                                                             float* __restrict__ out_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wg >= total_rows) return;
    const int64_t e = wg / rows_pad, r = wg % rows_pad;
    float* h = out + wg * k_pad;
    float* l = out_lo + wg * k_pad;
    if (r >= rows) {
        for (int64_t c = lane; c < k_pad; c += 32) {
            h[c] = 0.f;
            l[c] = 0.f;
        }
        return;
    }
    const float* src = x + (e * rows + r) * ld_x;
    // fast path: 16-byte aligned rows of at most 1024 floats stay in registers (one read of x)
    if (k <= 1024 && (k & 3) == 0 && (ld_x & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        constexpr int MAXV = 1024 / 128;  // float4 per lane
        float4 v[MAXV];
        const int nvec = (int)(k >> 2);
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = i * 32 + lane;
            v[i] = (c < nvec) ? __ldg(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += (double)v[i].x * v[i].x + (double)v[i].y * v[i].y + (double)v[i].z * v[i].z + (double)v[i].w * v[i].w;
        }
        float denom = 1.f;
        if (normalize) denom = fmaxf((float)sqrt(warp_sum(ss)), 1e-12f);
        const int nvec_pad = (int)(k_pad >> 2);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = i * 32 + lane;
            if (c >= nvec_pad) continue;
            float4 a = v[i];
            if (normalize) a = make_float4(__fdiv_rn(a.x, denom), __fdiv_rn(a.y, denom), __fdiv_rn(a.z, denom), __fdiv_rn(a.w, denom));
            reinterpret_cast<float4*>(h)[c] = a;
            reinterpret_cast<float4*>(l)[c] =
                make_float4(tf32_residual(a.x), tf32_residual(a.y), tf32_residual(a.z), tf32_residual(a.w));
        }
        return;
    }
    float denom = 1.f;
    if (normalize) {
        double ss = 0.0;
        for (int64_t c = lane; c < k; c += 32) {
            const double v = (double)src[c];
            ss += v * v;
        }
        ss = warp_sum(ss);
        denom = fmaxf((float)sqrt(ss), 1e-12f);
    }
    for (int64_t c = lane; c < k_pad; c += 32) {
        float v = 0.f;
        if (c < k) v = normalize ? __fdiv_rn(src[c], denom) : src[c];
        h[c] = v;
        l[c] = tf32_residual(v);
    }
}

// --------------------------------------------------------------------------------------------
// SIMT validation contraction: plain fp32 FFMA.  128 x 64 tile, 256 threads, 8 x 4 outputs per thread.
// Not the product path for throughput; it pins the tensor-core kernel.
// --------------------------------------------------------------------------------------------
constexpr int SIMT_BN = 64;
constexpr int SIMT_BK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmOperand A, GemmOperand B, int64_t K, GemmEpilogue ep) {
    __shared__ float sA[SIMT_BK][GEMM_BM + 4];
    __shared__ float sB[SIMT_BK][SIMT_BN + 4];
    __shared__ float sTile[GEMM_BM][SIMT_BN + 1];
    __shared__ unsigned char sFlags[GEMM_BM];
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int64_t e = blockIdx.z;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads: rows ty + 16*i (i<8), cols tx + 16*j (j<4)
    const int64_t ra0 = (int64_t)tile_m * GEMM_BM, rb0 = (int64_t)tile_n * SIMT_BN;
    const float* Ap = A.p + e * A.ep_stride;
    const float* Bp = B.p + e * B.ep_stride;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = 0; k0 < K; k0 += SIMT_BK) {
        // A: 128 rows x 16 k = 2048 values, 8 per thread; B: 64 x 16 = 1024, 4 per thread
        for (int i = tid; i < GEMM_BM * SIMT_BK; i += 256) {
            const int r = i / SIMT_BK, c = i % SIMT_BK;
            sA[c][r] = (ra0 + r < A.rows && k0 + c < K) ? Ap[(ra0 + r) * A.ld + k0 + c] : 0.f;
        }
        for (int i = tid; i < SIMT_BN * SIMT_BK; i += 256) {
            const int r = i / SIMT_BK, c = i % SIMT_BK;
            sB[c][r] = (rb0 + r < B.rows && k0 + c < K) ? Bp[(rb0 + r) * B.ld + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SIMT_BK; ++kk) {
            float a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sTile[ty + 16 * i][tx + 16 * j] = acc[i][j];
    __syncthreads();
    tile_epilogue<SIMT_BN>(&sTile[0][0], SIMT_BN + 1, sFlags, ep, e, tile_m, tile_n, tid, 256);
}

int gemm_simt(const GemmOperand& a, const GemmOperand& b, int E, int64_t M, int64_t N, int64_t K, const GemmEpilogue& ep,
              cudaStream_t s) {
    dim3 grid((unsigned)ceil_div64(N, SIMT_BN), (unsigned)ceil_div64(M, GEMM_BM), E);
    gemm_simt_kernel<<<grid, 256, 0, s>>>(a, b, K, ep);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

// --------------------------------------------------------------------------------------------
// A3 finalisation: one block per episode.
// --------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

template <typename T, typename Op>
__device__ T fin_block_reduce(T v, Op op, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T r = scratch[0];
    for (int w = 1; w < FIN_THREADS / 32; ++w) r = op(r, scratch[w]);
    return r;
}

// in-place (x - min) / (1e-7 + max - min) over n values of one episode, float32 like the reference
__device__ void block_minmax_inplace(float* v, int64_t n, float* scratch) {
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += FIN_THREADS) {
        mn = fminf(mn, v[i]);
        mx = fmaxf(mx, v[i]);
    }
    mn = fin_block_reduce(mn, [](float a, float b) { return fminf(a, b); }, scratch);
    mx = fin_block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, scratch);
    const float den = (1e-7f + mx) - mn;
    for (int64_t i = threadIdx.x; i < n; i += FIN_THREADS) v[i] = __fdiv_rn(v[i] - mn, den);
}

__global__ void __launch_bounds__(FIN_THREADS) vva_finalize_kernel(const float* __restrict__ colstats,
                                                                   const uint8_t* __restrict__ row_fg, int64_t M,
                                                                   int64_t N, int tiles_m, float* __restrict__ out) {
    __shared__ float s_f[FIN_THREADS / 32];
    __shared__ int s_i[FIN_THREADS / 32];
    const int64_t e = blockIdx.x;
    int t = 0;
#pragma unroll 8
    for (int64_t m = threadIdx.x; m < M; m += FIN_THREADS) t += row_fg[e * M + m] ? 1 : 0;
    t = fin_block_reduce(t, [](int a, int b) { return a + b; }, s_i);
    const int64_t n_bg = M - t;
    float* o = out + e * N;
    const float* cs = colstats + e * tiles_m * 4 * N;
    for (int64_t n = threadIdx.x; n < N; n += FIN_THREADS) {
        float fg_max = -INFINITY, bg_max = -INFINITY;
        double fg_sum = 0.0, bg_sum = 0.0;
        for (int tm0 = 0; tm0 < tiles_m; tm0 += 4) {  // 16 independent loads in flight (one CTA per episode)
            float v[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = tm0 + q < tiles_m;
                const float* c = cs + (int64_t)(ok ? tm0 + q : 0) * 4 * N + n;
                v[q][0] = ok ? __ldg(c) : -INFINITY;
                v[q][1] = ok ? __ldg(c + N) : 0.f;
                v[q][2] = ok ? __ldg(c + 2 * N) : -INFINITY;
                v[q][3] = ok ? __ldg(c + 3 * N) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // same order as a plain loop over the tiles
                fg_max = fmaxf(fg_max, v[q][0]);
                fg_sum += (double)v[q][1];
                bg_max = fmaxf(bg_max, v[q][2]);
                bg_sum += (double)v[q][3];
            }
        }
        // mean * max in float32; with no fg row the reference raises - we emit NaN (0/0)
        float v = (float)(fg_sum / (double)t) * fg_max;
        if (t == 0) v = NAN;
        if (n_bg > 0) v -= (float)(bg_sum / (double)n_bg) * bg_max;
        o[n] = v;
    }
    __syncthreads();
    block_minmax_inplace(o, N, s_f);
}

__global__ void __launch_bounds__(FIN_THREADS) resize_minmax_kernel(const float* __restrict__ src, int gs, int gd,
                                                                    int apply_minmax, float* __restrict__ out) {
    __shared__ float s_f[FIN_THREADS / 32];
    const int64_t e = blockIdx.x;
    const float scale = (float)gs / (float)gd;  // ATen nearest: src = min(floor(dst * scale), in - 1)
    float* o = out + e * gd * gd;
    for (int i = threadIdx.x; i < gd * gd; i += FIN_THREADS) {
        const int y = i / gd, x = i % gd;
        const int sy = min((int)floorf(y * scale), gs - 1), sx = min((int)floorf(x * scale), gs - 1);
        o[i] = src[e * gs * gs + sy * gs + sx];
    }
    __syncthreads();
    if (apply_minmax) block_minmax_inplace(o, (int64_t)gd * gd, s_f);
}

__global__ void __launch_bounds__(FIN_THREADS) minmax_rows_kernel(float* v, int64_t n) {
    __shared__ float s_f[FIN_THREADS / 32];
    block_minmax_inplace(v + (int64_t)blockIdx.x * n, n, s_f);
}

int minmax_rows(float* v, int E, int64_t n, cudaStream_t s) {
    minmax_rows_kernel<<<E, FIN_THREADS, 0, s>>>(v, n);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

static __thread char g_err[512];
char* last_error_buffer() { return g_err; }

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int marsb200_version(void) { return 100; }
const char* marsb200_last_error(void) { return last_error_buffer(); }
int64_t marsb200_pad_rows(int64_t rows) { return ceil_div64(rows, GEMM_BM) * GEMM_BM; }
int64_t marsb200_pad_k(int64_t k) { return ceil_div64(k, GEMM_PAD_K) * GEMM_PAD_K; }

int marsb200_normalize_rows(const float* x, int64_t ld_x, int E, int64_t rows, int64_t k, int normalize, float* out,
                            float* out_lo, void* stream) {
    MARS_REQUIRE(x && out && out_lo, "null pointer");
    MARS_REQUIRE(E > 0 && rows > 0 && k > 0 && ld_x >= k, "shape");
    const int64_t rows_pad = marsb200_pad_rows(rows), k_pad = marsb200_pad_k(k);
    const int64_t total = (int64_t)E * rows_pad;
    normalize_rows_kernel<<<(unsigned)ceil_div64(total, 8), 256, 0, as_stream(stream)>>>(
        x, ld_x, rows, k, rows_pad, k_pad, normalize, total, out, out_lo);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_sim_contract(const float* a, const float* a_lo, const float* b, const float* b_lo, int E, int64_t M,
                          int64_t N, int64_t K, float* sim_out, float* cost_out, const uint8_t* row_fg,
                          float* colstats, int backend, void* stream) {
    MARS_REQUIRE(a && a_lo && b && b_lo, "null operand");
    MARS_REQUIRE(E > 0 && E <= 65535 && M > 0 && N > 0 && K > 0, "shape");
    MARS_REQUIRE((row_fg == nullptr) == (colstats == nullptr), "row_fg and colstats go together");
    MARS_REQUIRE(sim_out || cost_out || colstats, "no output requested");
    GemmEpilogue ep{};
    ep.out0 = sim_out;
    ep.out1 = cost_out;
    ep.maxwith = nullptr;
    ep.row_fg = row_fg;
    ep.colstats = colstats;
    ep.M = M;
    ep.N = N;
    ep.ld_out = N;
    ep.ld_max = 0;
    ep.tiles_m = (int)(marsb200_pad_rows(M) / GEMM_BM);
    const int64_t m_pad = marsb200_pad_rows(M), n_pad = marsb200_pad_rows(N), k_pad = marsb200_pad_k(K);
    const GemmOperand oa{a, a_lo, m_pad, k_pad, m_pad * k_pad}, ob{b, b_lo, n_pad, k_pad, n_pad * k_pad};
    if (backend == MARSB200_GEMM_SIMT) return gemm_simt(oa, ob, E, M, N, K, ep, as_stream(stream));
    if (backend == MARSB200_GEMM_TCGEN05) return gemm_tcgen05(oa, ob, E, M, N, K, ep, as_stream(stream));
    return fail(MARSB200_ERR_ARG, "%s: unknown backend %lld", "marsb200_sim_contract", backend);
}

int marsb200_vva_finalize(const float* colstats, const uint8_t* row_fg, int E, int64_t M, int64_t N, float* out,
                          void* stream) {
    MARS_REQUIRE(colstats && row_fg && out, "null pointer");
    MARS_REQUIRE(E > 0 && M > 0 && N > 0, "shape");
    const int tiles_m = (int)(marsb200_pad_rows(M) / GEMM_BM);
    vva_finalize_kernel<<<E, FIN_THREADS, 0, as_stream(stream)>>>(colstats, row_fg, M, N, tiles_m, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_resize_minmax(const float* src, int E, int gs, int gd, int apply_minmax, float* out, void* stream) {
    MARS_REQUIRE(src && out, "null pointer");
    MARS_REQUIRE(E > 0 && gs > 0 && gd > 0, "shape");
    resize_minmax_kernel<<<E, FIN_THREADS, 0, as_stream(stream)>>>(src, gs, gd, apply_minmax, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

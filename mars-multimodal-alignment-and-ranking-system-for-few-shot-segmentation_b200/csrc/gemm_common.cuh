// Shared pieces of the NT contraction C[m,n] = sum_k A[m,k] B[n,k] used for S = Fs Fq^T (A2/A3)
// and D D^T (A4): operand layout, the epilogue parameter block and the tile epilogue that both the
// SIMT validation kernel runs once its 128 x BN fp32 tile sits in shared memory (the tcgen05 kernel has its
// own warp-specialised version of the same arithmetic).
#pragma once
#include "common.cuh"

namespace marsb200 {

constexpr int GEMM_BM = 128;  // rows per tile; also the row granularity of `colstats`
constexpr int GEMM_PAD_K = 32;

struct GemmEpilogue {
    float* out0;            // S (or R = max(maxwith, acc)); [E, M, ld_out] or null
    float* out1;            // cost = (1 - S) / 2; [E, M, ld_out] or null
    const float* maxwith;   // [E, M, ld_max] or null
    const uint8_t* row_fg;  // [E, M] or null
    float* colstats;        // [E, tiles_m, 4, N] or null
    int64_t M, N;
    int64_t ld_out, ld_max;
    int tiles_m;
    int symmetric;          // C == C^T (A == B, M == N): only tiles with tile_n >= tile_m are computed and
                            // every off-diagonal tile is also written mirrored (tcgen05 back end only)
};

// tile: BM x BN fp32 values in shared memory with row stride `lds` (floats).  All `nthreads` threads
// of the CTA call this.  flags_smem: BM bytes of scratch.
template <int BN>
__device__ __forceinline__ void tile_epilogue(const float* tile, int lds, unsigned char* flags_smem,
                                              const GemmEpilogue& ep, int64_t e, int tile_m, int tile_n, int tid,
                                              int nthreads) {
    const int64_t m0 = (int64_t)tile_m * GEMM_BM;
    const int64_t n0 = (int64_t)tile_n * BN;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;

    if (ep.out0 || ep.out1) {
        // each warp owns a contiguous band of rows; 4 rows x BN/32 columns are loaded before any store so
        // that the `maxwith` reads of a band overlap instead of serialising on global latency
        constexpr int CH = BN / 32;
        constexpr int RU = 4;
        const int rows_per_warp = GEMM_BM / nwarps;
        for (int rb = warp * rows_per_warp; rb < (warp + 1) * rows_per_warp; rb += RU) {
            float v[RU][CH];
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int64_t m = m0 + rb + i;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int64_t n = n0 + c * 32 + lane;
                    float mw = -INFINITY;
                    if (ep.maxwith && m < ep.M && n < ep.N) mw = ep.maxwith[(e * ep.M + m) * ep.ld_max + n];
                    const float t = tile[(rb + i) * lds + c * 32 + lane];
                    v[i][c] = ep.maxwith ? fmaxf(t, mw) : t;
                }
            }
#pragma unroll
            for (int i = 0; i < RU; ++i) {
                const int64_t m = m0 + rb + i;
                if (m >= ep.M) continue;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int64_t n = n0 + c * 32 + lane;
                    if (n < ep.N) {
                        if (ep.out0) ep.out0[(e * ep.M + m) * ep.ld_out + n] = v[i][c];
                        if (ep.out1) ep.out1[(e * ep.M + m) * ep.ld_out + n] = (1.0f - v[i][c]) / 2.0f;
                    }
                }
            }
        }
        if (ep.symmetric && tile_m != tile_n) {
            // mirrored block: output row = tile column, output columns = tile rows (BN == GEMM_BM here)
            const int cols_per_warp = BN / nwarps;
            constexpr int RH = GEMM_BM / 32;
            for (int cb = warp * cols_per_warp; cb < (warp + 1) * cols_per_warp; cb += RU) {
                float v[RU][RH];
#pragma unroll
                for (int i = 0; i < RU; ++i) {
                    const int64_t orow = n0 + cb + i;
#pragma unroll
                    for (int c = 0; c < RH; ++c) {
                        const int64_t ocol = m0 + c * 32 + lane;
                        float mw = -INFINITY;
                        if (ep.maxwith && orow < ep.M && ocol < ep.N) mw = ep.maxwith[(e * ep.M + orow) * ep.ld_max + ocol];
                        const float t = tile[(c * 32 + lane) * lds + cb + i];
                        v[i][c] = ep.maxwith ? fmaxf(t, mw) : t;
                    }
                }
#pragma unroll
                for (int i = 0; i < RU; ++i) {
                    const int64_t orow = n0 + cb + i;
                    if (orow >= ep.M) continue;
#pragma unroll
                    for (int c = 0; c < RH; ++c) {
                        const int64_t ocol = m0 + c * 32 + lane;
                        if (ocol < ep.N) {
                            if (ep.out0) ep.out0[(e * ep.M + orow) * ep.ld_out + ocol] = v[i][c];
                            if (ep.out1) ep.out1[(e * ep.M + orow) * ep.ld_out + ocol] = (1.0f - v[i][c]) / 2.0f;
                        }
                    }
                }
            }
        }
    }
    if (ep.colstats) {
        for (int r = tid; r < GEMM_BM; r += nthreads) {
            const int64_t m = m0 + r;
            flags_smem[r] = (m < ep.M) ? (ep.row_fg[e * ep.M + m] ? 1 : 2) : 0;  // 1 fg, 2 bg, 0 padding
        }
        __syncthreads();
        for (int c = tid; c < BN; c += nthreads) {
            const int64_t n = n0 + c;
            if (n >= ep.N) continue;
            float fg_max = -INFINITY, bg_max = -INFINITY;
            double fg_sum = 0.0, bg_sum = 0.0;
            for (int r = 0; r < GEMM_BM; ++r) {
                const float v = tile[r * lds + c];
                const unsigned char f = flags_smem[r];
                if (f == 1) {
                    fg_max = fmaxf(fg_max, v);
                    fg_sum += (double)v;
                } else if (f == 2) {
                    bg_max = fmaxf(bg_max, v);
                    bg_sum += (double)v;
                }
            }
            float* cs = ep.colstats + ((e * ep.tiles_m + tile_m) * 4) * ep.N + n;
            cs[0] = fg_max;
            cs[ep.N] = (float)fg_sum;
            cs[2 * ep.N] = bg_max;
            cs[3 * ep.N] = (float)bg_sum;
        }
    }
}

// lo = x - (x with the low 13 mantissa bits cleared): exact in fp32.  kind::tf32 reads only the top 19 bits of an
// fp32 word, so (x, lo) is the hi/lo pair of the error-compensated product without storing hi separately.
__device__ __forceinline__ float tf32_residual(float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// One fp32 operand of the contraction: [E][rows][ld] with k contiguous, plus its tf32 residual array of the same
// layout (only the tensor-core back end reads it).  Rows >= `rows` and columns >= K read as zero (TMA
// out-of-bounds fill in the tensor-core kernel, explicit checks in the SIMT kernel).
struct GemmOperand {
    const float* p;
    const float* lo;    // tf32_residual(p), same layout
    int64_t rows;       // rows present per episode
    int64_t ld;         // row stride in floats (multiple of 4, >= K)
    int64_t ep_stride;  // floats between episodes
};

// back ends (gemm_simt in vva.cu, gemm_tcgen05 in gemm_tc.cu)
int gemm_simt(const GemmOperand& a, const GemmOperand& b, int E, int64_t M, int64_t N, int64_t K, const GemmEpilogue& ep,
              cudaStream_t s);
int gemm_tcgen05(const GemmOperand& a, const GemmOperand& b, int E, int64_t M, int64_t N, int64_t K,
                 const GemmEpilogue& ep, cudaStream_t s);

}  // namespace marsb200

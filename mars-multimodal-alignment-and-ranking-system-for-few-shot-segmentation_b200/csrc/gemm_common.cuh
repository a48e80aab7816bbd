// Shared pieces of the NT contraction C[m,n] = sum_k A[m,k] B[n,k] used for S = Fs Fq^T (A2/A3)
// and D D^T (A4): operand layout, the epilogue parameter block and the tile epilogue that both the
// tcgen05 kernel and the SIMT validation kernel run once their 128 x BN fp32 tile sits in shared
// memory.
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
        for (int r = warp; r < GEMM_BM; r += nwarps) {
            const int64_t m = m0 + r;
            if (m >= ep.M) break;
#pragma unroll
            for (int c = lane; c < BN; c += 32) {
                const int64_t n = n0 + c;
                if (n < ep.N) {
                    float v = tile[r * lds + c];
                    if (ep.maxwith) v = fmaxf(v, ep.maxwith[(e * ep.M + m) * ep.ld_max + n]);
                    if (ep.out0) ep.out0[(e * ep.M + m) * ep.ld_out + n] = v;
                    if (ep.out1) ep.out1[(e * ep.M + m) * ep.ld_out + n] = (1.0f - v) / 2.0f;
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

// back ends (gemm_simt in vva.cu, gemm_tcgen05 in gemm_tc.cu).  Operands hi/lo: [E, rows_pad, k_pad].
int gemm_simt(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int E, int64_t M, int64_t N,
              int64_t K, const GemmEpilogue& ep, cudaStream_t s);
int gemm_tcgen05(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int E, int64_t M,
                 int64_t N, int64_t K, const GemmEpilogue& ep, cudaStream_t s);

}  // namespace marsb200

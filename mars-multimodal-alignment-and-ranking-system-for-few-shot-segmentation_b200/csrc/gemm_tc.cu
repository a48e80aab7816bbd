// tcgen05 contraction C[m,n] = sum_k A[m,k] B[n,k] for fp32 operands with fp32-grade accuracy:
// each operand is pre-split into hi = tf32(x) and lo = x - hi, and every 128x128x32 k-block issues
// hi*hi + hi*lo + lo*hi as kind::tf32 MMAs (3xTF32 error compensation) accumulating in fp32 in TMEM.
// Operand tiles arrive by TMA (128-byte swizzle) through a 3-stage mbarrier ring; one thread issues the
// MMAs; the epilogue moves the accumulator TMEM -> registers -> shared memory and runs tile_epilogue
// (coalesced S / cost / max(D, .) stores, fused fg/bg column statistics).
//
// Roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (warp 1 owns the TMEM
// allocation), all four warps = epilogue (warp w reads TMEM lanes 32w..32w+31).
#include <algorithm>
#include <cstdlib>

#include "gemm_common.cuh"
#include "tc_common.cuh"

namespace marsb200 {

using namespace tc;

constexpr int TC_BN = 128;
constexpr int TC_BK = 32;                            // fp32 elements = one 128 B swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = GEMM_BM * TC_BK * 4;   // 16 KB per operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;    // A_hi, A_lo, B_hi, B_lo
constexpr int TC_EPI_LDS = TC_BN + 1;                // padded fp32 row of the epilogue tile
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers, flags*/;
static_assert(GEMM_BM * TC_EPI_LDS * 4 <= TC_STAGES * TC_STAGE_BYTES, "epilogue tile must fit in the stage ring");

__global__ void __launch_bounds__(128, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    int num_kb, GemmEpilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;
    // barrier block: full[STAGES], empty[STAGES], tmem_full, then the TMEM address and the row flags
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * TC_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + TC_STAGES * TC_STAGE_BYTES + 8 * (2 * TC_STAGES + 1));
    unsigned char* flags = base_ptr + TC_STAGES * TC_STAGE_BYTES + 128;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int e = blockIdx.z;
    if (ep.symmetric) {  // blockIdx.x enumerates the upper-triangular tile pairs row by row
        int t = blockIdx.x, tm = 0;
        while (t >= ep.tiles_m - tm) {
            t -= ep.tiles_m - tm;
            ++tm;
        }
        tile_m = tm;
        tile_n = tm + t;
    }

    if (tid == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TC_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t phase = (kb / TC_STAGES) & 1;
                mbar_wait(empty_bar(s), phase ^ 1);
                mbar_expect_tx(full_bar(s), TC_STAGE_BYTES);
                const uint32_t st = base + s * TC_STAGE_BYTES;
                const int k0 = kb * TC_BK, m0 = tile_m * GEMM_BM, n0 = tile_n * TC_BN;
                tma_load_3d(st, &map_a_hi, full_bar(s), k0, m0, e);
                tma_load_3d(st + TC_TILE_BYTES, &map_a_lo, full_bar(s), k0, m0, e);
                tma_load_3d(st + 2 * TC_TILE_BYTES, &map_b_hi, full_bar(s), k0, n0, e);
                tma_load_3d(st + 3 * TC_TILE_BYTES, &map_b_lo, full_bar(s), k0, n0, e);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: 4 k-steps of 8 per block, 3 products per k-step
            constexpr uint32_t idesc = make_idesc(/*C=F32*/ 1, /*A=TF32*/ 2, /*B=TF32*/ 2, GEMM_BM, TC_BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t phase = (kb / TC_STAGES) & 1;
                mbar_wait(full_bar(s), phase);
                tc_fence_after();
                const uint32_t st = base + s * TC_STAGE_BYTES;
                const uint64_t a_hi = make_sw128_kmajor_desc(st);
                const uint64_t a_lo = make_sw128_kmajor_desc(st + TC_TILE_BYTES);
                const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * TC_TILE_BYTES);
                const uint64_t b_lo = make_sw128_kmajor_desc(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 B per k-step, in 16 B units
                    mma_tf32(tmem_acc, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                    mma_tf32(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1);
                    mma_tf32(tmem_acc, a_hi + adv, b_hi + adv, idesc, 1);
                }
                tc_commit(empty_bar(s));  // frees the stage once these MMAs have read it
            }
            tc_commit(tmem_full_bar);  // accumulator complete
        }
        __syncwarp();
    }

    // ---- epilogue: TMEM -> registers -> shared memory tile (reuses the stage ring)
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* tile = reinterpret_cast<float*>(base_ptr);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[row * TC_EPI_LDS + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, TC_BN);
    tile_epilogue<TC_BN>(tile, TC_EPI_LDS, flags, ep, e, tile_m, tile_n, tid, 128);
}

// ================================================================================================
// Persistent variant: one CTA per SM loops over tiles; the accumulator is double buffered in TMEM
// (2 x 128 columns) so that four dedicated epilogue warps drain tile i while the MMA thread already
// works on tile i+1 and the TMA thread prefetches operands across tile boundaries.
// Roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (+ TMEM alloc),
// warps 2..5 = epilogue (warp w drains TMEM lanes 32*(w%4).., one 32x32 chunk at a time through a
// private padded staging tile: coalesced stores, mirrored stores, per-warp column partials).
// ================================================================================================
constexpr int PG_THREADS = 192;
constexpr int PG_EPI_WARPS = 4;
constexpr int PG_STAGING_FLOATS = 32 * 33;                         // per epilogue warp
constexpr int PG_OFF_STAGING = TC_STAGES * TC_STAGE_BYTES;
constexpr int PG_OFF_PARTIAL = PG_OFF_STAGING + PG_EPI_WARPS * PG_STAGING_FLOATS * 4;
constexpr int PG_OFF_FLAGS = PG_OFF_PARTIAL + PG_EPI_WARPS * 4 * TC_BN * 4;   // [warp][stat][col] floats
constexpr int PG_OFF_BARS = PG_OFF_FLAGS + 128;
constexpr int PG_SMEM_BYTES = PG_OFF_BARS + 256 + 1024;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                               int num_kb, int tiles_n, int tiles_per_ep, int total_tiles, GemmEpilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + PG_OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    auto acc_full_bar = [&](int a) { return bars + 8u * (2 * TC_STAGES + a); };
    auto acc_empty_bar = [&](int a) { return bars + 8u * (2 * TC_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + PG_OFF_BARS + 8 * (2 * TC_STAGES + 4));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    auto decode = [&](int tile, int& e, int& tile_m, int& tile_n) {
        e = tile / tiles_per_ep;
        int t = tile - e * tiles_per_ep;
        if (ep.symmetric) {  // upper-triangular pairs, row by row
            int tm = 0;
            while (t >= ep.tiles_m - tm) {
                t -= ep.tiles_m - tm;
                ++tm;
            }
            tile_m = tm;
            tile_n = tm + t;
        } else {
            tile_m = t / tiles_n;
            tile_n = t - tile_m * tiles_n;
        }
    };

    if (tid == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full_bar(a), 1);
            mbar_init(acc_empty_bar(a), PG_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * TC_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int e, tile_m, tile_n;
                decode(tile, e, tile_m, tile_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t phase = (it / TC_STAGES) & 1;
                    mbar_wait(empty_bar(s), phase ^ 1);
                    mbar_expect_tx(full_bar(s), TC_STAGE_BYTES);
                    const uint32_t st = base + s * TC_STAGE_BYTES;
                    const int k0 = kb * TC_BK, m0 = tile_m * GEMM_BM, n0 = tile_n * TC_BN;
                    tma_load_3d(st, &map_a_hi, full_bar(s), k0, m0, e);
                    tma_load_3d(st + TC_TILE_BYTES, &map_a_lo, full_bar(s), k0, m0, e);
                    tma_load_3d(st + 2 * TC_TILE_BYTES, &map_b_hi, full_bar(s), k0, n0, e);
                    tma_load_3d(st + 3 * TC_TILE_BYTES, &map_b_lo, full_bar(s), k0, n0, e);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(/*C=F32*/ 1, /*A=TF32*/ 2, /*B=TF32*/ 2, GEMM_BM, TC_BN);
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                const uint32_t a = tl & 1, aphase = (tl >> 1) & 1;
                mbar_wait(acc_empty_bar(a), aphase ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + a * TC_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t phase = (it / TC_STAGES) & 1;
                    mbar_wait(full_bar(s), phase);
                    tc_fence_after();
                    const uint32_t st = base + s * TC_STAGE_BYTES;
                    const uint64_t a_hi = make_sw128_kmajor_desc(st);
                    const uint64_t a_lo = make_sw128_kmajor_desc(st + TC_TILE_BYTES);
                    const uint64_t b_hi = make_sw128_kmajor_desc(st + 2 * TC_TILE_BYTES);
                    const uint64_t b_lo = make_sw128_kmajor_desc(st + 3 * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                        mma_tf32(tmem_acc, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                        mma_tf32(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1);
                        mma_tf32(tmem_acc, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    tc_commit(empty_bar(s));
                }
                tc_commit(acc_full_bar(a));
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps
        const int q = warp & 3;   // TMEM lane quadrant this warp may read
        const int ew = warp - 2;  // 0..3: staging / partial slot
        float* staging = reinterpret_cast<float*>(base_ptr + PG_OFF_STAGING) + ew * PG_STAGING_FLOATS;
        float* partial = reinterpret_cast<float*>(base_ptr + PG_OFF_PARTIAL);  // [4 warps][4 stats][TC_BN]
        unsigned char* flags = base_ptr + PG_OFF_FLAGS + q * 32;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            int e, tile_m, tile_n;
            decode(tile, e, tile_m, tile_n);
            const uint32_t a = tl & 1, aphase = (tl >> 1) & 1;
            const int64_t m0 = (int64_t)tile_m * GEMM_BM + q * 32;  // first row of this warp's band
            const int64_t n0 = (int64_t)tile_n * TC_BN;
            if (ep.colstats) {
                const int64_t m = m0 + lane;
                flags[lane] = (m < ep.M) ? (ep.row_fg[(int64_t)e * ep.M + m] ? 1 : 2) : 0;
            }
            mbar_wait(acc_full_bar(a), aphase);
            tc_fence_after();
            const bool mirror = ep.symmetric && tile_m != tile_n;
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + a * TC_BN + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                if (c == TC_BN / 32 - 1) {  // accumulator fully read: hand it back to the MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty_bar(a));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) staging[lane * 33 + j] = __uint_as_float(v[j]);
                __syncwarp();
                const int64_t nc = n0 + c * 32;
                if (ep.out0 || ep.out1) {
                    // direct block: lane = column; all 32 rows of `maxwith` loads in flight at once
#pragma unroll 1
                    for (int r0 = 0; r0 < 32; r0 += 32) {
                        float val[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int64_t m = m0 + r0 + i, n = nc + lane;
                            float t = staging[(r0 + i) * 33 + lane];
                            if (ep.maxwith && m < ep.M && n < ep.N)
                                t = fmaxf(t, ep.maxwith[((int64_t)e * ep.M + m) * ep.ld_max + n]);
                            val[i] = t;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int64_t m = m0 + r0 + i, n = nc + lane;
                            if (m < ep.M && n < ep.N) {
                                if (ep.out0) ep.out0[((int64_t)e * ep.M + m) * ep.ld_out + n] = val[i];
                                if (ep.out1) ep.out1[((int64_t)e * ep.M + m) * ep.ld_out + n] = (1.0f - val[i]) / 2.0f;
                            }
                        }
                    }
                    if (mirror) {
                        // mirrored block: output row = tile column nc + cc, lane = tile row (contiguous in the output)
#pragma unroll 1
                        for (int c0 = 0; c0 < 32; c0 += 32) {
                            float val[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int64_t orow = nc + c0 + i, ocol = m0 + lane;
                                float t = staging[lane * 33 + c0 + i];
                                if (ep.maxwith && orow < ep.M && ocol < ep.N)
                                    t = fmaxf(t, ep.maxwith[((int64_t)e * ep.M + orow) * ep.ld_max + ocol]);
                                val[i] = t;
                            }
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int64_t orow = nc + c0 + i, ocol = m0 + lane;
                                if (orow < ep.M && ocol < ep.N) {
                                    if (ep.out0) ep.out0[((int64_t)e * ep.M + orow) * ep.ld_out + ocol] = val[i];
                                    if (ep.out1) ep.out1[((int64_t)e * ep.M + orow) * ep.ld_out + ocol] = (1.0f - val[i]) / 2.0f;
                                }
                            }
                        }
                    }
                }
                if (ep.colstats) {
                    // this warp's 32 rows of column nc + lane
                    float fg_max = -INFINITY, bg_max = -INFINITY;
                    double fg_sum = 0.0, bg_sum = 0.0;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const float t = staging[r * 33 + lane];
                        const unsigned char f = flags[r];
                        if (f == 1) {
                            fg_max = fmaxf(fg_max, t);
                            fg_sum += (double)t;
                        } else if (f == 2) {
                            bg_max = fmaxf(bg_max, t);
                            bg_sum += (double)t;
                        }
                    }
                    float* pw = partial + (ew * 4) * TC_BN + c * 32 + lane;
                    pw[0] = fg_max;
                    pw[TC_BN] = (float)fg_sum;
                    pw[2 * TC_BN] = bg_max;
                    pw[3 * TC_BN] = (float)bg_sum;
                }
                __syncwarp();  // staging is reused by the next chunk
            }
            if (ep.colstats) {
                epi_bar_sync();  // all four bands have written their partials
                const int col = (warp - 2) * 32 + lane;
                const int64_t n = n0 + col;
                if (n < ep.N) {
                    // combine the four 32-row bands in row order (TMEM quadrant order), deterministically
                    float fg_max = -INFINITY, bg_max = -INFINITY;
                    double fg_sum = 0.0, bg_sum = 0.0;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int w = (qq + 2) & 3;  // epilogue slot of the warp that owns quadrant qq (warp%4 == qq)
                        const float* pr = partial + (w * 4) * TC_BN + col;
                        fg_max = fmaxf(fg_max, pr[0]);
                        fg_sum += (double)pr[TC_BN];
                        bg_max = fmaxf(bg_max, pr[2 * TC_BN]);
                        bg_sum += (double)pr[3 * TC_BN];
                    }
                    float* cs = ep.colstats + (((int64_t)e * ep.tiles_m + tile_m) * 4) * ep.N + n;
                    cs[0] = fg_max;
                    cs[ep.N] = (float)fg_sum;
                    cs[2 * ep.N] = bg_max;
                    cs[3 * ep.N] = (float)bg_sum;
                }
                epi_bar_sync();  // partials may be overwritten by the next tile
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * TC_BN);
}

static int make_operand_map(CUtensorMap* map, const float* ptr, int E, int64_t rows_pad, int64_t k_pad, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(MARSB200_ERR_CUDA, "%s: cuTensorMapEncodeTiled entry point unavailable", "gemm_tcgen05");
    cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)rows_pad, (cuuint64_t)E};
    cuuint64_t strides[2] = {(cuuint64_t)k_pad * 4, (cuuint64_t)rows_pad * k_pad * 4};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MARSB200_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%lld)", "gemm_tcgen05", r);
    return MARSB200_OK;
}

int gemm_tcgen05(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int E, int64_t M,
                 int64_t N, int64_t K, const GemmEpilogue& ep, cudaStream_t s) {
    const int64_t m_pad = marsb200_pad_rows(M), n_pad = marsb200_pad_rows(N), k_pad = marsb200_pad_k(K);
    for (const float* p : {a_hi, a_lo, b_hi, b_lo})
        if (reinterpret_cast<uintptr_t>(p) & 15)
            return fail(MARSB200_ERR_ARG, "%s: operands must be 16-byte aligned", "gemm_tcgen05");
    CUtensorMap maps[4];
    int rc;
    if ((rc = make_operand_map(&maps[0], a_hi, E, m_pad, k_pad, GEMM_BM))) return rc;
    if ((rc = make_operand_map(&maps[1], a_lo, E, m_pad, k_pad, GEMM_BM))) return rc;
    if ((rc = make_operand_map(&maps[2], b_hi, E, n_pad, k_pad, TC_BN))) return rc;
    if ((rc = make_operand_map(&maps[3], b_lo, E, n_pad, k_pad, TC_BN))) return rc;
    static bool attr_set = false;
    static int num_sms = 148;
    static bool one_tile_per_cta = false;
    if (!attr_set) {
        MARS_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        MARS_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES));
        int dev = 0;
        MARS_CUDA_OK(cudaGetDevice(&dev));
        MARS_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        const char* env = getenv("MARSB200_GEMM_ONE_TILE");  // debugging: the non-persistent kernel
        one_tile_per_cta = env && env[0] == '1';
        attr_set = true;
    }
    const int tiles_m = (int)(m_pad / GEMM_BM), tiles_n = (int)(n_pad / TC_BN);
    int tiles_per_ep = tiles_m * tiles_n;
    if (ep.symmetric) {
        if (M != N || a_hi != b_hi || a_lo != b_lo)
            return fail(MARSB200_ERR_ARG, "%s: symmetric epilogue needs A == B", "gemm_tcgen05");
        tiles_per_ep = tiles_m * (tiles_m + 1) / 2;
    }
    if (one_tile_per_cta) {
        dim3 grid(ep.symmetric ? tiles_per_ep : tiles_n, ep.symmetric ? 1 : tiles_m, E);
        gemm_tcgen05_kernel<<<grid, 128, TC_SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], (int)(k_pad / TC_BK), ep);
    } else {
        const int total = tiles_per_ep * E;
        const int grid = std::min(total, num_sms);
        gemm_tcgen05_persistent_kernel<<<grid, PG_THREADS, PG_SMEM_BYTES, s>>>(
            maps[0], maps[1], maps[2], maps[3], (int)(k_pad / TC_BK), tiles_n, tiles_per_ep, total, ep);
    }
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // namespace marsb200

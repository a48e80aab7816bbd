// tcgen05 contraction C[m,n] = sum_k A[m,k] B[n,k] for fp32 operands with fp32-grade accuracy (3xTF32):
// x = hi + lo with hi = the top 19 bits of x (what kind::tf32 reads of an fp32 word: the tensor core ignores
// the low 13 mantissa bits, so the fp32 array itself is the `hi` operand) and lo = x - hi (exact, produced by the
// kernels that write the operands).  Per k-step of 8 the products hi*hi, hi*lo and lo*hi accumulate in fp32 in
// TMEM as TWO instructions: a_hi x [b_hi; b_lo] (N = 256: the b_lo tile sits right behind the b_hi tile in shared
// memory, so one descriptor covers both and the two products land in two 128-column halves of the accumulator)
// and a_lo x b_hi (N = 128).  That reads 20 KB of shared memory per k-step instead of 24 KB; the epilogue adds
// the two halves.
//
// TMA (128-byte swizzle) feeds a 3-stage ring of {A_hi, A_lo, B_hi, B_lo} 16 KB tiles.  Persistent: one CTA per
// SM loops over output tiles; the accumulator (256 columns) is double buffered in TMEM (all 512 columns), so the
// epilogue warps drain tile i while the MMA thread works on tile i+1 and the TMA thread prefetches operands
// across tile boundaries.  The issue of a tcgen05.mma blocks until the tensor pipe accepts it, so the MMA thread
// does nothing else per k-block than one barrier wait and one commit.
// Roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (+ TMEM alloc),
// warps 2..5 = epilogue (warp w drains TMEM lanes 32*(w%4).., one 32x32 chunk at a time through a private
// padded staging tile: coalesced stores, mirrored stores, per-warp column partials).
#include <algorithm>

#include "gemm_common.cuh"
#include "tc_common.cuh"

namespace marsb200 {

using namespace tc;

constexpr int TC_BN = 128;
constexpr int TC_BK = 32;                            // fp32 elements = one 128 B swizzle row
constexpr int TC_TILE_BYTES = GEMM_BM * TC_BK * 4;   // 16 KB per operand tile
constexpr int TC_PAIR_BYTES = 2 * TC_TILE_BYTES;     // A tile then B tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo (B_lo directly behind B_hi)
constexpr int TC_STAGES = 3;
constexpr int TC_ACC_COLS = 2 * TC_BN;             // [hi*hi + lo*hi | hi*lo]

constexpr int PG_THREADS = 192;
constexpr int PG_EPI_WARPS = 4;
constexpr int PG_STAGING_FLOATS = 32 * 33;                         // per epilogue warp
constexpr int PG_OFF_STAGING = TC_STAGES * TC_STAGE_BYTES;
constexpr int PG_OFF_PARTIAL = PG_OFF_STAGING + PG_EPI_WARPS * PG_STAGING_FLOATS * 4;
constexpr int PG_OFF_FLAGS = PG_OFF_PARTIAL + PG_EPI_WARPS * 4 * TC_BN * 4;   // [warp][stat][col] floats
constexpr int PG_OFF_BARS = PG_OFF_FLAGS + 128;
constexpr int PG_SMEM_BYTES = PG_OFF_BARS + 256 + 1024;
static_assert(PG_SMEM_BYTES <= 227 * 1024, "shared memory budget");

// -DMARSB200_GEMM_PROFILE: clock64 stamps of CTA 0 (MMA thread, epilogue warp 0) per tile, read back with
// marsb200_debug_gemm_profile (profiles/gemm_timeline.py).  Compiled out otherwise.
#ifdef MARSB200_GEMM_PROFILE
__device__ long long g_gemm_prof[128];
#define GP_STAMP(slot) do { if (blockIdx.x == 0) g_gemm_prof[slot] = clock64(); } while (0)
#define GP_CLOCK() clock64()
#define GP_PUT(slot, v) do { if (blockIdx.x == 0) g_gemm_prof[slot] = (v); } while (0)
#else
#define GP_STAMP(slot) do { } while (0)
#define GP_CLOCK() 0ll
#define GP_PUT(slot, v) do { } while (0)
#endif

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                               int num_kb, int tiles_n, int tiles_per_ep, int total_tiles, GemmEpilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + PG_OFF_BARS;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    auto acc_full = [&](int a) { return bars + 8u * (2 * TC_STAGES + a); };
    auto acc_empty = [&](int a) { return bars + 8u * (2 * TC_STAGES + 2 + a); };
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
            mbar_init(acc_full(a), 1);
            mbar_init(acc_empty(a), PG_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * TC_ACC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int e, tile_m, tile_n;
                decode(tile, e, tile_m, tile_n);
                const int m0 = tile_m * GEMM_BM, n0 = tile_n * TC_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(empty_bar(s), ((it / TC_STAGES) & 1) ^ 1);
                    mbar_expect_tx(full_bar(s), TC_STAGE_BYTES);
                    const uint32_t st = base + s * TC_STAGE_BYTES;
                    tma_load_3d(st, &map_a_hi, full_bar(s), kb * TC_BK, m0, e);
                    tma_load_3d(st + 2 * TC_TILE_BYTES, &map_b_hi, full_bar(s), kb * TC_BK, n0, e);
                    tma_load_3d(st + TC_TILE_BYTES, &map_a_lo, full_bar(s), kb * TC_BK, m0, e);
                    tma_load_3d(st + 3 * TC_TILE_BYTES, &map_b_lo, full_bar(s), kb * TC_BK, n0, e);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer: 4 k-steps of 8 per block, two instructions per k-step
            constexpr uint32_t idesc_n256 = make_idesc(/*C=F32*/ 1, /*A=TF32*/ 2, /*B=TF32*/ 2, GEMM_BM, 2 * TC_BN);
            constexpr uint32_t idesc_n128 = make_idesc(/*C=F32*/ 1, /*A=TF32*/ 2, /*B=TF32*/ 2, GEMM_BM, TC_BN);
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                const uint32_t a = tl & 1, aphase = (tl >> 1) & 1;
                GP_STAMP(tl * 8 + 0);
                mbar_wait(acc_empty(a), aphase ^ 1);  // the epilogue has drained this accumulator
                GP_STAMP(tl * 8 + 1);
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + a * TC_ACC_COLS;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(full_bar(s), (it / TC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = base + s * TC_STAGE_BYTES;
                    const uint64_t a_hi = make_sw128_kmajor_desc(st);
                    const uint64_t a_lo = make_sw128_kmajor_desc(st + TC_TILE_BYTES);
                    const uint64_t b_hl = make_sw128_kmajor_desc(st + 2 * TC_TILE_BYTES);  // 256 rows: b_hi then b_lo
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 B per k-step, in 16 B units
                        mma_tf32(tmem_acc, a_hi + adv, b_hl + adv, idesc_n256, (kb | k) != 0);
                        mma_tf32(tmem_acc, a_lo + adv, b_hl + adv, idesc_n128, 1);
                    }
                    tc_commit(empty_bar(s));  // the stage is free once these MMAs have read it
                }
                tc_commit(acc_full(a));
                GP_STAMP(tl * 8 + 2);
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
            if (ew == 0 && lane == 0) GP_STAMP(tl * 8 + 3);
            mbar_wait(acc_full(a), aphase);
            if (ew == 0 && lane == 0) GP_STAMP(tl * 8 + 4);
            tc_fence_after();
            const bool mirror = ep.symmetric && tile_m != tile_n;
            float* o0 = ep.out0 ? ep.out0 + (int64_t)e * ep.M * ep.ld_out : nullptr;
            float* o1 = ep.out1 ? ep.out1 + (int64_t)e * ep.M * ep.ld_out : nullptr;
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t v[32], w[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * TC_ACC_COLS + (uint32_t)(c * 32);
                tmem_ld_32x32(taddr, v);
                tmem_ld_32x32(taddr + TC_BN, w);
                tmem_ld_wait();
                if (c == TC_BN / 32 - 1) {  // accumulator fully read: hand it back to the MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(a));
                    if (ew == 0 && lane == 0) GP_STAMP(tl * 8 + 5);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) staging[lane * 33 + j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
                __syncwarp();
                const int64_t nc = n0 + c * 32;
                if (o0 || o1) {
                    // direct block: lane = column, 32 coalesced row segments
                    const int64_t n = nc + lane;
                    if (n < ep.N) {
#pragma unroll 8
                        for (int i = 0; i < 32; ++i) {
                            const int64_t m = m0 + i;
                            if (m < ep.M) {
                                const float t = staging[i * 33 + lane];
                                if (o0) o0[m * ep.ld_out + n] = t;
                                if (o1) o1[m * ep.ld_out + n] = (1.0f - t) / 2.0f;
                            }
                        }
                    }
                    if (mirror) {
                        // mirrored block: output row = tile column nc + i, lane = tile row (contiguous in the output)
                        const int64_t ocol = m0 + lane;
                        if (ocol < ep.N) {
#pragma unroll 8
                            for (int i = 0; i < 32; ++i) {
                                const int64_t orow = nc + i;
                                if (orow < ep.M) {
                                    const float t = staging[lane * 33 + i];
                                    if (o0) o0[orow * ep.ld_out + ocol] = t;
                                    if (o1) o1[orow * ep.ld_out + ocol] = (1.0f - t) / 2.0f;
                                }
                            }
                        }
                    }
                }
                if (ep.colstats) {
                    // this warp's 32 rows of column nc + lane; fp32 sums over 32 values (four interleaved chains),
                    // everything above the band is combined in double.  (fp64 arithmetic here ran ~6x slower
                    // whenever the MMA thread was issuing.)
                    float fg_max = -INFINITY, bg_max = -INFINITY;
                    float fs[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float t = staging[r * 33 + lane];
                        const unsigned char f = flags[r];
                        const bool is_fg = (f == 1), is_bg = (f == 2);
                        fg_max = is_fg ? fmaxf(fg_max, t) : fg_max;
                        bg_max = is_bg ? fmaxf(bg_max, t) : bg_max;
                        fs[r & 3] += is_fg ? t : 0.f;
                        bs[r & 3] += is_bg ? t : 0.f;
                    }
                    const float fg_sum = (fs[0] + fs[1]) + (fs[2] + fs[3]);
                    const float bg_sum = (bs[0] + bs[1]) + (bs[2] + bs[3]);
                    float* pw = partial + (ew * 4) * TC_BN + c * 32 + lane;
                    pw[0] = fg_max;
                    pw[TC_BN] = fg_sum;
                    pw[2 * TC_BN] = bg_max;
                    pw[3 * TC_BN] = bg_sum;
                }
                __syncwarp();  // staging is reused by the next chunk
            }
            if (ew == 0 && lane == 0) GP_STAMP(tl * 8 + 7);
            if (ep.colstats) {
                epi_bar_sync();  // all four bands have written their partials
                if (ew == 0 && lane == 0) GP_STAMP(64 + tl * 2);
                const int col = ew * 32 + lane;
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
                if (ew == 0 && lane == 0) GP_STAMP(64 + tl * 2 + 1);
                epi_bar_sync();  // partials may be overwritten by the next tile
            }
            if (ew == 0 && lane == 0) GP_STAMP(tl * 8 + 6);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * TC_ACC_COLS);
}

// 3-D tensor map [K, rows, E] of an fp32 array; elements outside (rows, K) arrive as zeros
static int make_operand_map(CUtensorMap* map, const float* ptr, const GemmOperand& op, int E, int64_t K) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(MARSB200_ERR_CUDA, "%s: cuTensorMapEncodeTiled entry point unavailable", "gemm_tcgen05");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)op.rows, (cuuint64_t)E};
    cuuint64_t strides[2] = {(cuuint64_t)op.ld * 4, (cuuint64_t)op.ep_stride * 4};
    cuuint32_t box[3] = {TC_BK, GEMM_BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MARSB200_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%lld)", "gemm_tcgen05", r);
    return MARSB200_OK;
}

#ifdef MARSB200_GEMM_PROFILE
}  // namespace marsb200
extern "C" int marsb200_debug_gemm_profile(long long* out128, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out128, marsb200::g_gemm_prof, sizeof(long long) * 128);
    if (reset) {
        long long z[128] = {0};
        cudaMemcpyToSymbol(marsb200::g_gemm_prof, z, sizeof(z));
    }
    return 0;
}
namespace marsb200 {
#endif

int gemm_tcgen05(const GemmOperand& a, const GemmOperand& b, int E, int64_t M, int64_t N, int64_t K,
                 const GemmEpilogue& ep, cudaStream_t s) {
    for (const GemmOperand* o : {&a, &b})
        if (!o->lo || ((reinterpret_cast<uintptr_t>(o->p) | reinterpret_cast<uintptr_t>(o->lo)) & 15) || (o->ld & 3) ||
            (o->ep_stride & 3) || o->ld < K)
            return fail(MARSB200_ERR_ARG,
                        "%s: operands need their residual array, 16-byte alignment and strides that are multiples of 4 floats",
                        "gemm_tcgen05");
    CUtensorMap maps[4];
    int rc;
    if ((rc = make_operand_map(&maps[0], a.p, a, E, K))) return rc;
    if ((rc = make_operand_map(&maps[1], a.lo, a, E, K))) return rc;
    if ((rc = make_operand_map(&maps[2], b.p, b, E, K))) return rc;
    if ((rc = make_operand_map(&maps[3], b.lo, b, E, K))) return rc;
    static PerDeviceOnce configured;  // the attribute is per device
    MARS_CUDA_OK(per_device_once(configured, [] {
        return cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM_BYTES);
    }));
    int num_sms = 0;  // the stream's partition when it belongs to a green context
    if ((rc = sms_for_stream(s, &num_sms))) return rc;
    const int tiles_m = (int)ceil_div64(M, GEMM_BM), tiles_n = (int)ceil_div64(N, TC_BN);
    int tiles_per_ep = tiles_m * tiles_n;
    if (ep.symmetric) {
        if (M != N || a.p != b.p) return fail(MARSB200_ERR_ARG, "%s: symmetric epilogue needs A == B", "gemm_tcgen05");
        tiles_per_ep = tiles_m * (tiles_m + 1) / 2;
    }
    const int total = tiles_per_ep * E;
    const int grid = std::min(total, num_sms);
    gemm_tcgen05_persistent_kernel<<<grid, PG_THREADS, PG_SMEM_BYTES, s>>>(
        maps[0], maps[1], maps[2], maps[3], (int)ceil_div64(K, TC_BK), tiles_n, tiles_per_ep, total, ep);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // namespace marsb200

// tcgen05 contraction C[m,n] = sum_k A[m,k] B[n,k] for fp32 operands with fp32-grade accuracy:
// each operand is pre-split into hi = tf32(x) and lo = x - hi, and every 128x128x32 k-block issues
// hi*hi + hi*lo + lo*hi as kind::tf32 MMAs (3xTF32 error compensation) accumulating in fp32 in TMEM.
// Operand tiles arrive by TMA (128-byte swizzle) through a 3-stage mbarrier ring; one thread issues the
// MMAs; the epilogue moves the accumulator TMEM -> registers -> shared memory and runs tile_epilogue
// (coalesced S / cost / max(D, .) stores, fused fg/bg column statistics).
//
// Roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (warp 1 owns the TMEM
// allocation), all four warps = epilogue (warp w reads TMEM lanes 32w..32w+31).
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
    if (!attr_set) {
        MARS_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((unsigned)(n_pad / TC_BN), (unsigned)(m_pad / GEMM_BM), E);
    if (ep.symmetric) {
        if (M != N || a_hi != b_hi || a_lo != b_lo)
            return fail(MARSB200_ERR_ARG, "%s: symmetric epilogue needs A == B", "gemm_tcgen05");
        const unsigned t = (unsigned)(m_pad / GEMM_BM);
        grid = dim3(t * (t + 1) / 2, 1, E);
    }
    gemm_tcgen05_kernel<<<grid, 128, TC_SMEM_BYTES, s>>>(maps[0], maps[1], maps[2], maps[3], (int)(k_pad / TC_BK), ep);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // namespace marsb200

// One pass over the proposal masks at HBM rate that produces BOTH the packed bits and the pairwise
// intersection matrix (P <= 256): the float32 masks are read exactly once (streaming 128-bit loads),
// each 128-pixel slice of a row is turned into (a) 4 packed words written back to HBM and (b) one
// 128-byte SWIZZLE_128B shared-memory row of u8 0/1, and a single thread issues tcgen05.mma kind::i8
// on the staged tiles while the next slices stream in.  The tensor work (68.7 G MACs per c2 episode)
// hides completely under the 1.07 GB read.
//
// CTA = one pixel slice of one episode, all P rows: 16 producer warps (warp w owns rows 16w..16w+15,
// a warp reads one row slice per load: 512 contiguous bytes), 1 MMA warp, 6-stage mbarrier ring of
// 32 KB tiles, two 128x256 s32 accumulators (all 512 TMEM columns), integer-atomic split-K reduction.
#include <algorithm>

#include "tc_common.cuh"

namespace marsb200 {

using namespace tc;

constexpr int FZ_ROWS = 256;
constexpr int FZ_TILE_BYTES = FZ_ROWS * 128;   // 32 KB: 256 rows x 128 pixels as u8
constexpr int FZ_STAGES = 6;
constexpr int FZ_PRODUCER_WARPS = 16;
constexpr int FZ_ROWS_PER_WARP = FZ_ROWS / FZ_PRODUCER_WARPS;  // 16
constexpr int FZ_BATCH = 8;                                    // rows in flight per warp
constexpr int FZ_THREADS = (FZ_PRODUCER_WARPS + 1) * 32;
constexpr int FZ_SMEM_BYTES = FZ_STAGES * FZ_TILE_BYTES + 1024 + 256;

// FULL: P == 256 and HW a multiple of 1024 -> no bounds checks anywhere in the producer loop
template <bool FULL>
__global__ void __launch_bounds__(FZ_THREADS, 1)
pack_pairwise_f32_kernel(const float* __restrict__ masks, int P, int64_t HW, int64_t wpm, int kb_per_split,
                         uint32_t* __restrict__ bits, int32_t* __restrict__ inter) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + FZ_STAGES * FZ_TILE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (FZ_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * FZ_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + FZ_STAGES * FZ_TILE_BYTES + 8 * (2 * FZ_STAGES + 1));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t e = blockIdx.y;
    const int total_kb = (int)(wpm / 4);
    const int kb_begin = blockIdx.x * kb_per_split;
    const int kb_end = min(kb_begin + kb_per_split, total_kb);
    const int num_kb = kb_end - kb_begin;
    const int m_tiles = (P + 127) / 128;

    // rows >= P are never written by the producers: clear the whole ring once
    for (int i = tid; i < FZ_STAGES * FZ_TILE_BYTES / 16; i += FZ_THREADS)
        reinterpret_cast<uint4*>(base_ptr)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int s = 0; s < FZ_STAGES; ++s) {
            mbar_init(full_bar(s), FZ_PRODUCER_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == FZ_PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp < FZ_PRODUCER_WARPS) {
        const int row0 = warp * FZ_ROWS_PER_WARP;
        const float* ebase = masks + e * (int64_t)P * HW;
        uint32_t* ebits = bits + e * (int64_t)P * wpm;
        static_assert(FZ_ROWS_PER_WARP == 2 * FZ_BATCH, "two register batches per k-block");

        // per-lane cursors: row `row0`, this lane's 4 pixels of k-block `kb_begin`; rows are HW floats apart
        const float* src = ebase + (int64_t)row0 * HW + (int64_t)kb_begin * 128 + lane * 4;
        uint32_t* dst = ebits + (int64_t)row0 * wpm + (int64_t)kb_begin * 4 + (lane >> 3);
        // byte offset of this lane's 4 operand bytes inside row r of a stage: r*128 + ((lane/4 ^ r%8) * 16) + lane%4*4;
        // row0 is a multiple of 16, so r%8 == (b0 + b) % 8
        const uint32_t lane_word = (uint32_t)(lane & 3) << 2;

        // issue the 8 row-slice loads of one half k-block (512 contiguous bytes per row, one row per load)
        auto issue = [&](int i, int b0, uint4 (&v)[FZ_BATCH]) {
#pragma unroll
            for (int b = 0; b < FZ_BATCH; ++b) {
                const float* p = src + (int64_t)(b0 + b) * HW + (int64_t)i * 128;
                if (FULL) {
                    v[b] = ldg_stream_u4(p);
                } else {
                    const int r = row0 + b0 + b;
                    const int64_t px = (int64_t)(kb_begin + i) * 128 + lane * 4;
                    v[b] = (r < P && px < HW) ? ldg_stream_u4(p) : make_uint4(0, 0, 0, 0);
                }
            }
        };
        // turn 8 loaded row slices into u8 operand rows (shared memory) and packed words (global)
        auto process = [&](int i, int b0, unsigned char* st, const uint4 (&v)[FZ_BATCH]) {
#pragma unroll
            for (int b = 0; b < FZ_BATCH; ++b) {
                const int rl = b0 + b;  // row within the warp's band
                const uint32_t nib = (__uint_as_float(v[b].x) > 0.f ? 1u : 0u) | (__uint_as_float(v[b].y) > 0.f ? 2u : 0u) |
                                     (__uint_as_float(v[b].z) > 0.f ? 4u : 0u) | (__uint_as_float(v[b].w) > 0.f ? 8u : 0u);
                const uint32_t off = (uint32_t)(row0 + rl) * 128u + ((((uint32_t)(lane >> 2)) ^ (uint32_t)(rl & 7)) << 4) + lane_word;
                uint32_t w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(0xffffffffu, w, 1);
                w |= __shfl_xor_sync(0xffffffffu, w, 2);
                w |= __shfl_xor_sync(0xffffffffu, w, 4);
                if (FULL || row0 + rl < P) {
                    *reinterpret_cast<uint32_t*>(st + off) = (nib * 0x00204081u) & 0x01010101u;
                    if ((lane & 7) == 0) dst[(int64_t)rl * wpm + (int64_t)i * 4] = w;
                }
            }
        };

        // software pipeline: the next half's loads are always in flight while the current half is processed
        uint4 va[FZ_BATCH], vb[FZ_BATCH];
        if (num_kb > 0) issue(0, 0, va);
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % FZ_STAGES;
            const uint32_t phase = (i / FZ_STAGES) & 1;
            unsigned char* st = base_ptr + s * FZ_TILE_BYTES;
            issue(i, FZ_BATCH, vb);
            mbar_wait(empty_bar(s), phase ^ 1);
            process(i, 0, st, va);
            if (i + 1 < num_kb) issue(i + 1, 0, va);
            process(i, FZ_BATCH, st, vb);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
        }
    } else {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(/*C=S32*/ 2, /*A=u8*/ 0, /*B=u8*/ 0, 128, FZ_ROWS);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % FZ_STAGES;
                const uint32_t phase = (i / FZ_STAGES) & 1;
                mbar_wait(full_bar(s), phase);
                tc_fence_after();
                const uint32_t st = base + s * FZ_TILE_BYTES;
                const uint64_t b_desc = make_sw128_kmajor_desc(st);
                for (int mt = 0; mt < m_tiles; ++mt) {
                    const uint64_t a_desc = make_sw128_kmajor_desc(st + mt * 128 * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        mma_i8(tmem_acc + mt * FZ_ROWS, a_desc + adv, b_desc + adv, idesc, (i | k) != 0);
                    }
                }
                tc_commit(empty_bar(s));
            }
            tc_commit(tmem_full_bar);
        }
        __syncwarp();
    }

    // epilogue: 16 warps drain TMEM (warp w: lane quadrant w%4, 64-column quarter w/4)
    if (num_kb > 0 && warp < FZ_PRODUCER_WARPS) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quad = warp & 3, quarter = warp >> 2;
        int32_t* out = inter + e * (int64_t)P * P;
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int i = mt * 128 + quad * 32 + lane;
            for (int c = 0; c < 2; ++c) {
                const int col0 = quarter * 64 + c * 32;
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * FZ_ROWS + col0), v);
                tmem_ld_wait();
                if (i < P) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int j = col0 + q;
                        const int val = (int)v[q];
                        if (j < P && val != 0) atomicAdd(&out[(int64_t)i * P + j], val);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FZ_PRODUCER_WARPS) tmem_dealloc(tmem_acc, 512);
}

}  // namespace marsb200

using namespace marsb200;

extern "C" int marsb200_pack_pairwise(const void* masks, int mask_dtype, int E, int P, int64_t HW, uint32_t* bits,
                                      int32_t* inter, int pair_backend, void* stream) {
    MARS_REQUIRE(masks && bits && inter, "null pointer");
    MARS_REQUIRE(E > 0 && E <= 65535 && P > 0 && HW > 0, "shape");
    const int64_t wpm = marsb200_words_per_mask(HW);
    cudaStream_t s = as_stream(stream);
    const bool fusable = pair_backend == MARSB200_PAIR_MMA && mask_dtype == MARSB200_MASK_F32 && P <= FZ_ROWS &&
                         HW % 4 == 0 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0;
    if (!fusable) {  // two-kernel path: any dtype / alignment / P
        int rc = marsb200_pack_masks(masks, mask_dtype, (int64_t)E * P, HW, bits, stream);
        if (rc) return rc;
        return marsb200_pairwise_inter(bits, E, P, wpm, inter, pair_backend, stream);
    }
    const int total_kb = (int)(wpm / 4);
    // all CTAs resident at once (1 per SM): split the pixels so that E * ksplit <= 148
    int device_sms = 148;
    MARS_CUDA_OK(device_sm_count(&device_sms));
    int ksplit = std::max(1, std::min(total_kb, device_sms / E));
    int kb_per_split = ceil_div(total_kb, ksplit);
    ksplit = ceil_div(total_kb, kb_per_split);
    static PerDeviceOnce attr_set;  // the attribute is per device
    MARS_CUDA_OK(per_device_once(attr_set, [] {
        cudaError_t e = cudaFuncSetAttribute(pack_pairwise_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES);
        return e != cudaSuccess ? e : cudaFuncSetAttribute(pack_pairwise_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES);
    }));
    MARS_CUDA_OK(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)E * P * P, s));
    if (P == FZ_ROWS && HW % 1024 == 0)
        pack_pairwise_f32_kernel<true><<<dim3(ksplit, E), FZ_THREADS, FZ_SMEM_BYTES, s>>>((const float*)masks, P, HW, wpm,
                                                                                        kb_per_split, bits, inter);
    else
        pack_pairwise_f32_kernel<false><<<dim3(ksplit, E), FZ_THREADS, FZ_SMEM_BYTES, s>>>((const float*)masks, P, HW, wpm,
                                                                                         kb_per_split, bits, inter);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

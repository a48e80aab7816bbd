// Pairwise intersections on the tensor cores: inter = Mb Mb^T with Mb the 0/1 mask matrix.
//
// The packed bits stay the HBM/L2 format (1 bit per pixel); producer warps expand 128 pixels of a
// row into one 128-byte swizzled shared-memory row of uint8 0/1 (the K-major SWIZZLE_128B operand
// layout), and one thread issues tcgen05.mma kind::i8 (u8 x u8 -> s32, exact) into TMEM.
// A CTA owns one pair (I, J) of 256-row blocks with I <= J and a slice of the pixels:
// two 128x256 accumulators = all 512 TMEM columns.  On the diagonal (I == J) the A operand is a
// sub-range of the B tile, so each row is expanded once.  Slices are reduced with integer atomics
// (exact, order independent).
#include <algorithm>

#include "tc_common.cuh"

namespace marsb200 {

using namespace tc;

constexpr int PM_ROWS = 256;                    // rows per block (B operand = MMA N)
constexpr int PM_KB_PIX = 128;                  // pixels per k-block = one 128 B swizzle row of u8
constexpr int PM_TILE_BYTES = PM_ROWS * 128;    // 32 KB per 256-row operand tile
constexpr int PM_STAGE_BYTES = 2 * PM_TILE_BYTES;  // B tile, then the A tile (off-diagonal pairs only)
constexpr int PM_STAGES = 3;
constexpr int PM_PRODUCER_WARPS = 8;
constexpr int PM_THREADS = (PM_PRODUCER_WARPS + 1) * 32;
constexpr int PM_SMEM_BYTES = PM_STAGES * PM_STAGE_BYTES + 1024 + 256;

// 4 mask bits -> 4 bytes of 0/1
__device__ __forceinline__ uint32_t spread_nibble(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// one row of a k-block: 128 bits (uint4) -> 8 swizzled 16-byte chunks
__device__ __forceinline__ void expand_row(unsigned char* tile, int row, uint4 bits) {
    const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
    unsigned char* dst = tile + row * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {  // chunk c = pixels 16c .. 16c+15 = half of word c/2
        const uint32_t half = (w[c >> 1] >> ((c & 1) * 16)) & 0xffffu;
        uint4 v;
        v.x = spread_nibble(half & 0xfu);
        v.y = spread_nibble((half >> 4) & 0xfu);
        v.z = spread_nibble((half >> 8) & 0xfu);
        v.w = spread_nibble((half >> 12) & 0xfu);
        *reinterpret_cast<uint4*>(dst + ((c ^ (row & 7)) << 4)) = v;
    }
}

__global__ void __launch_bounds__(PM_THREADS, 1)
pairwise_mma_kernel(const uint32_t* __restrict__ bits, int P, int64_t wpm, int blocks, int total_kb, int kb_per_split,
                    int32_t* __restrict__ inter) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + PM_STAGES * PM_STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (PM_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * PM_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + PM_STAGES * PM_STAGE_BYTES + 8 * (2 * PM_STAGES + 1));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // blockIdx.x -> upper-triangular block pair (bi <= bj); blockIdx.y -> pixel slice; blockIdx.z -> episode
    int t = blockIdx.x, bi = 0;
    while (t >= blocks - bi) {
        t -= blocks - bi;
        ++bi;
    }
    const int bj = bi + t;
    const bool diagonal = (bi == bj);
    const int64_t e = blockIdx.z;
    // total_kb: k-blocks (4 words = 128 pixels) of the pixel slice this launch covers, starting at `bits`
    const int kb_begin = blockIdx.y * kb_per_split;
    const int kb_end = min(kb_begin + kb_per_split, total_kb);
    const int num_kb = kb_end - kb_begin;
    const int rows_a = min(PM_ROWS, P - bi * PM_ROWS);  // valid rows of the A block
    const int m_tiles = (rows_a + 127) / 128;

    if (tid == 0) {
        for (int s = 0; s < PM_STAGES; ++s) {
            mbar_init(full_bar(s), PM_PRODUCER_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == PM_PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    const uint32_t* ebits = bits + e * P * wpm;

    if (warp < PM_PRODUCER_WARPS) {
        // ---- producers: thread `tid` owns row tid of the B block (and of the A block off the diagonal)
        const int rb = bj * PM_ROWS + tid, ra = bi * PM_ROWS + tid;
        const uint4* src_b = reinterpret_cast<const uint4*>(ebits + (int64_t)rb * wpm);
        const uint4* src_a = reinterpret_cast<const uint4*>(ebits + (int64_t)ra * wpm);
        // the words of k-block i + 2 are requested before k-block i is expanded: with the bits coming from HBM (the
        // packed masks of a launch exceed the L2) one load latency per k-block would otherwise pace the MMAs
        constexpr int AHEAD = 2;
        uint4 qb[AHEAD], qa[AHEAD];
#pragma unroll
        for (int a = 0; a < AHEAD; ++a) {
            qb[a] = (rb < P && a < num_kb) ? __ldg(src_b + kb_begin + a) : make_uint4(0, 0, 0, 0);
            qa[a] = (!diagonal && ra < P && a < num_kb) ? __ldg(src_a + kb_begin + a) : make_uint4(0, 0, 0, 0);
        }
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % PM_STAGES;
            const uint32_t phase = (i / PM_STAGES) & 1;
            const uint4 vb = qb[0], va = qa[0];
#pragma unroll
            for (int a = 0; a + 1 < AHEAD; ++a) {
                qb[a] = qb[a + 1];
                qa[a] = qa[a + 1];
            }
            const int nk = kb_begin + i + AHEAD;
            qb[AHEAD - 1] = (rb < P && i + AHEAD < num_kb) ? __ldg(src_b + nk) : make_uint4(0, 0, 0, 0);
            qa[AHEAD - 1] = (!diagonal && ra < P && i + AHEAD < num_kb) ? __ldg(src_a + nk) : make_uint4(0, 0, 0, 0);
            mbar_wait(empty_bar(s), phase ^ 1);
            unsigned char* st = base_ptr + s * PM_STAGE_BYTES;
            expand_row(st, tid, vb);
            if (!diagonal) expand_row(st + PM_TILE_BYTES, tid, va);
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
        }
    } else {
        // ---- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(/*C=S32*/ 2, /*A=u8*/ 0, /*B=u8*/ 0, 128, PM_ROWS);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % PM_STAGES;
                const uint32_t phase = (i / PM_STAGES) & 1;
                mbar_wait(full_bar(s), phase);
                tc_fence_after();
                const uint32_t st = base + s * PM_STAGE_BYTES;
                const uint64_t b_desc = make_sw128_kmajor_desc(st);
                const uint32_t a_base = diagonal ? st : st + PM_TILE_BYTES;
                for (int mt = 0; mt < m_tiles; ++mt) {
                    const uint64_t a_desc = make_sw128_kmajor_desc(a_base + mt * 128 * 128);
                    if (diagonal && mt == 1) {
                        // symmetric block: rows 128..255 only need columns 128..255 (N = 128); the lower-left
                        // quarter is the mirror of what m-tile 0 computed
                        constexpr uint32_t idesc_half = make_idesc(2, 0, 0, 128, 128);
                        const uint64_t b_half = make_sw128_kmajor_desc(st + 128 * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            mma_i8(tmem_acc + PM_ROWS + 128, a_desc + adv, b_half + adv, idesc_half, (i | k) != 0);
                        }
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 32 bytes (= 32 u8 elements) per MMA
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        mma_i8(tmem_acc + mt * PM_ROWS, a_desc + adv, b_desc + adv, idesc, (i | k) != 0);
                    }
                }
                tc_commit(empty_bar(s));
            }
            tc_commit(tmem_full_bar);
        }
        __syncwarp();
    }

    // ---- epilogue: warps 0..7 drain TMEM (warp w: lane quadrant w%4, column half w/4) with integer atomics
    if (num_kb > 0 && warp < PM_PRODUCER_WARPS) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        int32_t* out = inter + e * (int64_t)P * P;
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int i = bi * PM_ROWS + mt * 128 + quad * 32 + lane;
            for (int c = 0; c < 4; ++c) {
                const int col0 = half * 128 + c * 32;
                if (diagonal && mt == 1 && col0 < 128) continue;  // not computed: mirrored from m-tile 0 below
                const bool mirror = !diagonal || (mt == 0 && col0 >= 128);
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * PM_ROWS + col0), v);
                tmem_ld_wait();
                if (i < P) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int j = bj * PM_ROWS + col0 + q;
                        const int val = (int)v[q];
                        if (j < P && val != 0) {
                            atomicAdd(&out[(int64_t)i * P + j], val);
                            if (mirror) atomicAdd(&out[(int64_t)j * P + i], val);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PM_PRODUCER_WARPS) tmem_dealloc(tmem_acc, 512);
}

int pairwise_mma(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s, int64_t word_begin,
                 int64_t word_count, bool accumulate) {
    if (reinterpret_cast<uintptr_t>(bits) & 15)
        return fail(MARSB200_ERR_ARG, "%s: packed masks must be 16-byte aligned", "pairwise_mma");
    if (word_count < 0) word_count = wpm - word_begin;
    if (word_begin < 0 || word_begin % 4 || word_count <= 0 || word_count % 4 || word_begin + word_count > wpm)
        return fail(MARSB200_ERR_ARG, "%s: pixel slice must be whole 128-pixel blocks inside the mask (%lld, %lld)", "pairwise_mma",
                    word_begin, word_count);
    const int blocks = ceil_div(P, PM_ROWS);
    const int pairs = blocks * (blocks + 1) / 2;
    const int total_kb = (int)(word_count / 4);
    static PerDeviceOnce configured;  // the attribute is per device
    MARS_CUDA_OK(per_device_once(configured, [] {
        return cudaFuncSetAttribute(pairwise_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PM_SMEM_BYTES);
    }));
    int num_sms = 0;  // the stream's partition when it belongs to a green context
    if (int rc = sms_for_stream(s, &num_sms)) return rc;
    // one CTA per SM (all of TMEM, 193 KB of shared memory): split the pixels so that the launch fills whole waves
    // from below - 160 CTAs on 148 SMs would run two waves with the second one 8 % full
    const int units = pairs * E;
    const int waves = ceil_div(units, num_sms);
    int ksplit = std::max(1, std::min(total_kb, waves * num_sms / units));
    int kb_per_split = ceil_div(total_kb, ksplit);
    ksplit = ceil_div(total_kb, kb_per_split);
    if (!accumulate) MARS_CUDA_OK(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)E * P * P, s));
    dim3 grid(pairs, ksplit, E);
    pairwise_mma_kernel<<<grid, PM_THREADS, PM_SMEM_BYTES, s>>>(bits + word_begin, P, wpm, blocks, total_kb, kb_per_split, inter);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // namespace marsb200

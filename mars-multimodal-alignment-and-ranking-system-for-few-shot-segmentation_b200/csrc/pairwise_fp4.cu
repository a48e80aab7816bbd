// Pairwise intersections of up to 256 proposals as a block-scaled FP4 product: inter = Mb Mb^T with the 0/1 mask
// matrix held as e2m1 nibbles (1.0 = 0b0010) and every scale factor 1.0 (UE8M0 0x7f).  tcgen05.mma kind::mxf4
// issues at twice the rate of kind::i8; the products are 0 or 1 and the fp32 accumulators hold integers below 2^24,
// so the counts are exact (checked bit for bit against the popcount kernel in the tests).
//
// Same structure as pairwise_mma_kernel (pairwise_tc.cu) for the diagonal block only: a CTA owns all rows of an
// episode and a slice of the pixels; two producer threads per row expand its 256 pixels into one 128-byte SWIZZLE_128B
// shared-memory row per k-block; one thread issues M=128,N=256 (rows 0..127 x all) and M=128,N=128 (rows 128..255 x
// columns 128..255, the rest is the mirror) per 64 pixels.  Any permutation of the pixels inside a k-block is applied
// to both operands alike, so the nibble order inside a byte does not matter.  All scale-factor bytes are equal, so
// neither does their TMEM layout: 32 columns of 0x7f7f7f7f serve as SFA and SFB.
#include <algorithm>

#include "tc_common.cuh"

namespace marsb200 {

using namespace tc;

constexpr int PF_ROWS = 256;
constexpr int PF_KB_PIX = 256;                 // pixels per k-block = one 128 B swizzle row of fp4
constexpr int PF_TILE_BYTES = PF_ROWS * 128;   // 32 KB
constexpr int PF_STAGES = 6;
constexpr int PF_PRODUCER_WARPS = 16;        // thread t expands half h = t / 256 (128 pixels) of row t % 256
constexpr int PF_EPILOGUE_WARPS = 8;
constexpr int PF_THREADS = (PF_PRODUCER_WARPS + 1) * 32;
constexpr int PF_SMEM_BYTES = PF_STAGES * PF_TILE_BYTES + 1024 + 256;
constexpr uint32_t PF_COL_ACC1 = 256;          // second accumulator (rows 128..255 x columns 128..255)
constexpr uint32_t PF_COL_SF = 384;            // 32 columns of scale factors

// 16 mask bits -> 16 e2m1 nibbles (two words).  The bits are spread so that nibble j of `x` holds bit pair j, then a
// byte permute looks every pair up in the table {00 -> 0x00, 01 -> 0x02, 10 -> 0x20, 11 -> 0x22} held in one register.
__device__ __forceinline__ void spread_half_fp4(uint32_t h, uint32_t& lo, uint32_t& hi) {
    uint32_t x = h;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    lo = __byte_perm(0x22200200u, 0u, x);        // pixels 0..7
    hi = __byte_perm(0x22200200u, 0u, x >> 16);  // pixels 8..15
}

// half h of a row's k-block: words 4h .. 4h+3 = pixels 128h .. 128h+127 -> chunks 4h .. 4h+3 of the 128-byte row
__device__ __forceinline__ void expand_half_row_fp4(unsigned char* tile, int row, int h, uint4 bits) {
    const uint32_t w[4] = {bits.x, bits.y, bits.z, bits.w};
    unsigned char* dst = tile + row * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // chunk = one word = 32 pixels -> 16 bytes
        uint4 v;
        spread_half_fp4(w[c] & 0xffffu, v.x, v.y);
        spread_half_fp4(w[c] >> 16, v.z, v.w);
        *reinterpret_cast<uint4*>(dst + (((4 * h + c) ^ (row & 7)) << 4)) = v;
    }
}

__device__ __forceinline__ void mma_mxf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate, uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

// instruction descriptor of the block-scaled kinds: [7,10) A format, [10,13) B format (kind::mxf4: 1 = E2M1),
// [17,23) N >> 3, bit 23 scale format (1 = UE8M0), [24,29) M >> 4, bit 31 K (0 = 64); scale-factor ids 0.
__host__ __device__ constexpr uint32_t make_idesc_mxf4(uint32_t m, uint32_t n) {
    return (1u << 7) | (1u << 10) | ((n >> 3) << 17) | (1u << 23) | ((m >> 4) << 24);
}

__global__ void __launch_bounds__(PF_THREADS, 1)
pairwise_fp4_kernel(const uint32_t* __restrict__ bits, int P, int64_t wpm, int total_kb, int kb_per_split,
                    int32_t* __restrict__ inter) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + PF_STAGES * PF_TILE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (PF_STAGES + s); };
    const uint32_t tmem_full_bar = bars + 8u * (2 * PF_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + PF_STAGES * PF_TILE_BYTES + 8 * (2 * PF_STAGES + 1));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t e = blockIdx.y;
    // total_kb: k-blocks (8 words = 256 pixels) of the pixel slice this launch covers, starting at `bits`
    const int kb_begin = blockIdx.x * kb_per_split;
    const int kb_end = min(kb_begin + kb_per_split, total_kb);
    const int num_kb = kb_end - kb_begin;
    const int m_tiles = (P + 127) / 128;

    if (tid == 0) {
        for (int s = 0; s < PF_STAGES; ++s) {
            mbar_init(full_bar(s), PF_PRODUCER_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == PF_PRODUCER_WARPS) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    if (warp < 4) {  // scale factors: every byte 0x7f = 2^0, written to all 128 lanes of 32 columns
        const uint32_t one = 0x7f7f7f7fu;
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + PF_COL_SF;
#pragma unroll
        for (int c = 0; c < 32; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(one)
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t* ebits = bits + e * P * wpm;

    if (warp < PF_PRODUCER_WARPS) {
        const int r = tid & (PF_ROWS - 1), h = tid >> 8;
        const uint4* src = reinterpret_cast<const uint4*>(ebits + (int64_t)r * wpm) + h;
        constexpr int AHEAD = 2;
        uint4 q[AHEAD];
#pragma unroll
        for (int a = 0; a < AHEAD; ++a)
            q[a] = (r < P && a < num_kb) ? __ldg(src + 2 * (kb_begin + a)) : make_uint4(0, 0, 0, 0);
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % PF_STAGES;
            const uint32_t phase = (i / PF_STAGES) & 1;
            const uint4 cur = q[0];
#pragma unroll
            for (int a = 0; a + 1 < AHEAD; ++a) q[a] = q[a + 1];
            q[AHEAD - 1] = (r < P && i + AHEAD < num_kb) ? __ldg(src + 2 * (kb_begin + i + AHEAD)) : make_uint4(0, 0, 0, 0);
            mbar_wait(empty_bar(s), phase ^ 1);
            expand_half_row_fp4(base_ptr + s * PF_TILE_BYTES, r, h, cur);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(s));
        }
    } else {
        if (lane == 0) {
            constexpr uint32_t idesc_full = make_idesc_mxf4(128, PF_ROWS);
            constexpr uint32_t idesc_half = make_idesc_mxf4(128, 128);
            const uint32_t sfa = tmem_acc + PF_COL_SF, sfb = tmem_acc + PF_COL_SF + 16;
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % PF_STAGES;
                const uint32_t phase = (i / PF_STAGES) & 1;
                mbar_wait(full_bar(s), phase);
                tc_fence_after();
                const uint32_t st = base + s * PF_TILE_BYTES;
                const uint64_t b_desc = make_sw128_kmajor_desc(st);
                const uint64_t b_half = make_sw128_kmajor_desc(st + 128 * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // 32 bytes = 64 fp4 elements per MMA
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    mma_mxf4(tmem_acc, b_desc + adv, b_desc + adv, idesc_full, (i | k) != 0, sfa, sfb);
                }
                if (m_tiles > 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        mma_mxf4(tmem_acc + PF_COL_ACC1, b_half + adv, b_half + adv, idesc_half, (i | k) != 0, sfa, sfb);
                    }
                }
                tc_commit(empty_bar(s));
            }
            tc_commit(tmem_full_bar);
        }
        __syncwarp();
    }

    // ---- epilogue: warps 0..7 drain TMEM (warp w: lane quadrant w % 4, column half w / 4), integer atomics
    if (num_kb > 0 && warp < PF_EPILOGUE_WARPS) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        int32_t* out = inter + e * (int64_t)P * P;
        for (int mt = 0; mt < m_tiles; ++mt) {
            const int i = mt * 128 + quad * 32 + lane;
            for (int c = 0; c < 4; ++c) {
                const int col0 = half * 128 + c * 32;
                if (mt == 1 && col0 < 128) continue;  // mirrored from m-tile 0
                const bool mirror = (mt == 0 && col0 >= 128);
                const uint32_t tcol = mt == 0 ? (uint32_t)col0 : PF_COL_ACC1 + (uint32_t)(col0 - 128);
                uint32_t v[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)(quad * 32) << 16) + tcol, v);
                tmem_ld_wait();
                if (i < P) {
#pragma unroll
                    for (int qq = 0; qq < 32; ++qq) {
                        const int j = col0 + qq;
                        const int val = __float2int_rn(__uint_as_float(v[qq]));
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
    if (warp == PF_PRODUCER_WARPS) tmem_dealloc(tmem_acc, 512);
}

int pairwise_fp4(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s, int64_t word_begin,
                 int64_t word_count, bool accumulate) {
    if (reinterpret_cast<uintptr_t>(bits) & 15)
        return fail(MARSB200_ERR_ARG, "%s: packed masks must be 16-byte aligned", "pairwise_fp4");
    if (P > PF_ROWS) return fail(MARSB200_ERR_UNSUPPORTED, "%s: at most 256 proposals (%lld given)", "pairwise_fp4", P);
    if (wpm * 32 >= (1ll << 24)) return fail(MARSB200_ERR_UNSUPPORTED, "%s: masks of 2^24 pixels or more", "pairwise_fp4");
    if (word_count < 0) word_count = wpm - word_begin;
    if (word_begin < 0 || word_begin % 8 || word_count <= 0 || word_count % 8 || word_begin + word_count > wpm)
        return fail(MARSB200_ERR_ARG, "%s: pixel slice must be whole 256-pixel blocks inside the mask (%lld, %lld)", "pairwise_fp4",
                    word_begin, word_count);
    const int total_kb = (int)(word_count / 8);
    static PerDeviceOnce configured;  // the attribute is per device
    MARS_CUDA_OK(per_device_once(configured, [] {
        return cudaFuncSetAttribute(pairwise_fp4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM_BYTES);
    }));
    int num_sms = 0;
    if (int rc = sms_for_stream(s, &num_sms)) return rc;
    const int waves = ceil_div(E, num_sms);
    int ksplit = std::max(1, std::min(total_kb, waves * num_sms / E));
    int kb_per_split = ceil_div(total_kb, ksplit);
    ksplit = ceil_div(total_kb, kb_per_split);
    if (!accumulate) MARS_CUDA_OK(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)E * P * P, s));
    pairwise_fp4_kernel<<<dim3(ksplit, E), PF_THREADS, PF_SMEM_BYTES, s>>>(bits + word_begin, P, wpm, total_kb, kb_per_split, inter);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // namespace marsb200

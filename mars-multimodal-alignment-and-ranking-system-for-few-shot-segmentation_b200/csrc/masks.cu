// Mask ingest and mask-geometry kernels: bit-pack, patch pooling, region sums, pairwise
// AND+popcount intersections, OR-merge, point lookups, evaluator areas.
//
// HBM layout: a packed proposal is one row of `wpm = words_per_mask(HW)` uint32 words, bit k of
// word w = pixel 32*w + k of the flattened [H,W] mask; wpm is a multiple of 32 words (128 B) and
// the tail is zero.  One episode's proposals are P consecutive rows.
#include <algorithm>

#include "common.cuh"

namespace marsb200 {

// --------------------------------------------------------------------------------------------
// pack: [n, HW] float32 / uint8  ->  [n, wpm] bits.  HBM-bound: reads 4 (or 1) B/pixel once with
// 128-bit streaming loads, writes 1/32 (1/8) of that.  One warp turns 4 x 512 B into 16 words.
// --------------------------------------------------------------------------------------------
constexpr int PACK_THREADS = 256;
constexpr int PACK_UNROLL = 4;
// float32 ingest keeps 8 x 16 bytes per thread in flight: the same 7.1 TB/s as 4 on the whole device, but 75 instead of
// 66 GB/s per SM when the launch only owns a green-context partition of the SMs (partition.py)
constexpr int PACK_UNROLL_F32 = 8;

// f32: a thread's float4 gives a nibble; 8 lanes make a word.
__global__ void __launch_bounds__(PACK_THREADS) pack_f32_vec_kernel(const float* __restrict__ masks, int64_t n,
                                                                     int64_t HW, int64_t wpm,
                                                                     uint32_t* __restrict__ bits, int chunks) {
    const int64_t blk = blockIdx.x;
    const int64_t m = blk / chunks;
    const int chunk = (int)(blk % chunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PX_PER_BLOCK = PACK_THREADS * PACK_UNROLL_F32 * 4;  // 8192 pixels = 256 words
    const int64_t px_base = (int64_t)chunk * PX_PER_BLOCK + (int64_t)warp * (PACK_UNROLL_F32 * 128);
    const float* src = masks + m * HW;

    uint4 v[PACK_UNROLL_F32];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        const int64_t px = px_base + u * 128 + lane * 4;
        v[u] = (px < HW) ? ldg_stream_u4(src + px) : make_uint4(0, 0, 0, 0);  // HW % 4 == 0 on this path
    }
    uint32_t word[PACK_UNROLL_F32];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        uint32_t nib = (__uint_as_float(v[u].x) > 0.f ? 1u : 0u) | (__uint_as_float(v[u].y) > 0.f ? 2u : 0u) |
                       (__uint_as_float(v[u].z) > 0.f ? 4u : 0u) | (__uint_as_float(v[u].w) > 0.f ? 8u : 0u);
        uint32_t w = nib << (4 * (lane & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        word[u] = w;  // every lane of octet o = lane>>3 holds word (u, o)
    }
    // lane t stores word (u = t>>2, o = t&3): one 128 B store per warp
    uint32_t out = 0;
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        uint32_t got = __shfl_sync(0xffffffffu, word[u], (lane & 3) * 8);
        if ((lane >> 2) == u) out = got;
    }
    const int64_t w_idx = px_base / 32 + lane;
    if (lane < PACK_UNROLL_F32 * 4 && w_idx < wpm) bits[m * wpm + w_idx] = out;
}

// u8: a thread's uint4 gives 16 bits; 2 lanes make a word.
__global__ void __launch_bounds__(PACK_THREADS) pack_u8_vec_kernel(const uint8_t* __restrict__ masks, int64_t n,
                                                                    int64_t HW, int64_t wpm,
                                                                    uint32_t* __restrict__ bits, int chunks) {
    const int64_t blk = blockIdx.x;
    const int64_t m = blk / chunks;
    const int chunk = (int)(blk % chunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PX_PER_BLOCK = PACK_THREADS * PACK_UNROLL * 16;  // 16384 pixels = 512 words
    const int64_t px_base = (int64_t)chunk * PX_PER_BLOCK + (int64_t)warp * (PACK_UNROLL * 512);
    const uint8_t* src = masks + m * HW;

    uint4 v[PACK_UNROLL];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
        const int64_t px = px_base + u * 512 + lane * 16;
        v[u] = (px < HW) ? ldg_stream_u4(src + px) : make_uint4(0, 0, 0, 0);  // HW % 16 == 0 on this path
    }
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
        auto nib = [](uint32_t x) -> uint32_t {
            // bytes > 0 -> 0xff, keep bit 0 of each byte, gather the four into bits 24..27
            uint32_t mbits = __vcmpgtu4(x, 0u) & 0x01010101u;
            return (mbits * 0x01020408u) >> 24 & 0xfu;
        };
        uint32_t half = nib(v[u].x) | (nib(v[u].y) << 4) | (nib(v[u].z) << 8) | (nib(v[u].w) << 12);
        uint32_t w = half << (16 * (lane & 1));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        const int64_t w_idx = px_base / 32 + u * 16 + (lane >> 1);
        if ((lane & 1) == 0 && w_idx < wpm) bits[m * wpm + w_idx] = w;
    }
}

// generic path (any HW / alignment): one pixel per lane, a ballot per word.
template <typename T>
__global__ void __launch_bounds__(256) pack_scalar_kernel(const T* __restrict__ masks, int64_t n, int64_t HW,
                                                          int64_t wpm, uint32_t* __restrict__ bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = n * wpm;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = warp_global; w < total; w += warps) {
        const int64_t m = w / wpm, wi = w % wpm;
        const int64_t px = wi * 32 + lane;
        bool on = false;
        if (px < HW) on = (float)masks[m * HW + px] > 0.f;
        const uint32_t word = __ballot_sync(0xffffffffu, on);
        if (lane == 0) bits[w] = word;
    }
}

// --------------------------------------------------------------------------------------------
// support-mask pooling: one block per (mask, patch row).  The block streams the image rows of that patch-row bin
// (28-29 rows at 1024 -> 37) with coalesced loads - a thread keeps the columns x = tid, tid + 256, ... and walks down the
// rows - and ORs "any pixel > 0" into the column bins of each of its columns.  (The first version took one warp per
// (mask, bin): 28-float row segments, 1 TB/s.)
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pool_mask_kernel(const T* __restrict__ masks, int64_t n, int H, int W, int g,
                                                        uint8_t* __restrict__ out) {
    __shared__ int s_hit[256];  // g <= 255 column bins
    const int64_t m = blockIdx.x / g;
    const int by = (int)(blockIdx.x % g);
    const int y0 = bin_start(by, H, g), y1 = bin_end(by, H, g);
    for (int c = threadIdx.x; c < g; c += blockDim.x) s_hit[c] = 0;
    __syncthreads();
    const T* base = masks + (m * H + y0) * (int64_t)W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        bool any = false;
        int y = y0;
        for (; y + 4 <= y1; y += 4) {  // four independent loads in flight
            const T a = base[(int64_t)(y - y0) * W + x], b = base[(int64_t)(y - y0 + 1) * W + x];
            const T c = base[(int64_t)(y - y0 + 2) * W + x], d = base[(int64_t)(y - y0 + 3) * W + x];
            any |= ((float)a > 0.f) | ((float)b > 0.f) | ((float)c > 0.f) | ((float)d > 0.f);
        }
        for (; y < y1; ++y) any |= (float)base[(int64_t)(y - y0) * W + x] > 0.f;
        if (any)
            for (int c = bin_lo_of(x, W, g); c <= bin_hi_of(x, W, g); ++c) s_hit[c] = 1;  // benign race: every writer stores 1
    }
    __syncthreads();
    for (int c = threadIdx.x; c < g; c += blockDim.x) out[(m * g + by) * g + c] = s_hit[c] ? 1 : 0;
}

// --------------------------------------------------------------------------------------------
// pooled bitmap / area / pooled count of a packed mask: one block per mask, the packed row is read
// once with 128-bit loads (4 in flight per thread).  Pass 1 ORs every non-zero word into a
// row-aligned bit row per patch-row bin (shared memory, g x (W/32 + 1) words); pass 2 tests the
// column window of every bin.  Bins follow the adaptive-pool rule and may overlap.
// --------------------------------------------------------------------------------------------
constexpr int POOL_THREADS = 256;
constexpr int POOL_UNROLL = 4;

__device__ __forceinline__ void smem_or(uint32_t* p, uint32_t v) {
    if ((*p & v) != v) atomicOr(p, v);
}

__global__ void __launch_bounds__(POOL_THREADS) pool_packed_kernel(const uint32_t* __restrict__ bits, int64_t n, int H,
                                                                   int W, int g, int64_t wpm, int npw, int rw,
                                                                   uint32_t* __restrict__ pooled,
                                                                   int32_t* __restrict__ area,
                                                                   int32_t* __restrict__ pooled_count) {
    extern __shared__ __align__(16) uint32_t s_mem_pool[];
    uint32_t* binrow = s_mem_pool;             // g * rw
    uint32_t* s_pool = binrow + g * rw;        // npw
    int* s_cnt = reinterpret_cast<int*>(s_pool + npw);  // 2
    uint16_t* s_rowbins = reinterpret_cast<uint16_t*>(s_cnt + 2);  // H: first | last << 8 patch-row bin of row y
    const int64_t m = blockIdx.x;
    for (int i = threadIdx.x; i < g * rw + npw + 2; i += POOL_THREADS) s_mem_pool[i] = 0;
    for (int y = threadIdx.x; y < H; y += POOL_THREADS)
        s_rowbins[y] = (uint16_t)(bin_lo_of32(y, H, g) | (bin_hi_of32(y, H, g) << 8));
    __syncthreads();

    const uint4* row = reinterpret_cast<const uint4*>(bits + m * wpm);
    const int quads = (int)(wpm / 4);
    const int quads_per_row = (W % 128 == 0) ? W / 128 : 0;  // > 0 selects the word-aligned fast path
    const int qpr_shift = (quads_per_row > 0 && (quads_per_row & (quads_per_row - 1)) == 0) ? __ffs(quads_per_row) - 1 : -1;
    int my_area = 0;
    for (int q0 = threadIdx.x; q0 < quads; q0 += POOL_THREADS * POOL_UNROLL) {
        uint4 v[POOL_UNROLL];
#pragma unroll
        for (int u = 0; u < POOL_UNROLL; ++u) {
            const int q = q0 + u * POOL_THREADS;
            v[u] = (q < quads) ? __ldg(row + q) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < POOL_UNROLL; ++u) {
            if ((v[u].x | v[u].y | v[u].z | v[u].w) == 0) continue;
            const uint32_t words[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            const uint32_t quad = (uint32_t)(q0 + u * POOL_THREADS);
            if (quads_per_row > 0) {
                // W % 128 == 0: the four words sit in one image row, word-aligned with the bit rows
                // quads_per_row is a power of two for the usual widths (1024 -> 8): a shift instead of a division
                const uint32_t y = qpr_shift >= 0 ? (quad >> qpr_shift) : quad / (uint32_t)quads_per_row;
                if ((int)y < H) {
                    const int wi0 = (int)(quad - y * (uint32_t)quads_per_row) * 4;
                    const int rb = s_rowbins[y];
                    my_area += __popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]);
                    // rw is a multiple of 4 on this path: one 128-bit read tells whether the bit row already holds
                    // these pixels (the usual case inside a blob: ~28 image rows share a patch-row bin)
                    for (int jy = rb & 0xff; jy <= (rb >> 8); ++jy) {
                        uint32_t* dst = &binrow[jy * rw + wi0];
                        const uint4 cur = *reinterpret_cast<const uint4*>(dst);
                        if (((cur.x & words[0]) ^ words[0]) | ((cur.y & words[1]) ^ words[1]) |
                            ((cur.z & words[2]) ^ words[2]) | ((cur.w & words[3]) ^ words[3])) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (words[k] & ~(&cur.x)[k]) atomicOr(dst + k, words[k]);
                        }
                    }
                }
                continue;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t word = words[k];
                if (word == 0) continue;
                my_area += __popc(word);
                const uint32_t px = (quad * 4u + k) * 32u;
                int y = (int)(px / (uint32_t)W);
                int x = (int)(px - (uint32_t)y * (uint32_t)W);
                int left = 32;
                while (word != 0 && y < H) {  // a word may straddle rows when W % 32 != 0
                    const int len = min(left, W - x);
                    const uint32_t seg = (len == 32) ? word : (word & ((1u << len) - 1u));
                    if (seg != 0) {
                        const int wi = x >> 5, sh = x & 31;
                        const uint32_t lo = seg << sh;
                        const uint32_t hi = sh ? (seg >> (32 - sh)) : 0u;
                        const int rb = s_rowbins[y];
                        const int jy0 = rb & 0xff, jy1 = rb >> 8;
                        for (int jy = jy0; jy <= jy1; ++jy) {
                            if (lo) smem_or(&binrow[jy * rw + wi], lo);
                            if (hi) smem_or(&binrow[jy * rw + wi + 1], hi);
                        }
                    }
                    word = (len == 32) ? 0u : (word >> len);
                    left -= len;
                    x = 0;
                    ++y;
                }
            }
        }
    }
    my_area = warp_sum(my_area);
    if ((threadIdx.x & 31) == 0 && my_area) atomicAdd(&s_cnt[0], my_area);
    __syncthreads();
    for (int b = threadIdx.x; b < g * g; b += POOL_THREADS) {
        const int jy = b / g, jx = b - jy * g;
        const int xs = bin_start32(jx, W, g), xe = bin_end32(jx, W, g);  // [xs, xe)
        bool hit = false;
        for (int wi = xs >> 5; wi <= (xe - 1) >> 5; ++wi) {
            const int lo = max(xs - wi * 32, 0), hi = min(xe - wi * 32, 32);
            const uint32_t window = ((hi - lo) == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            hit |= (binrow[jy * rw + wi] & window) != 0;
        }
        if (hit) atomicOr(&s_pool[b >> 5], 1u << (b & 31));
    }
    __syncthreads();
    int pc = 0;
    for (int i = threadIdx.x; i < npw; i += POOL_THREADS) {
        const uint32_t v = s_pool[i];
        pooled[m * npw + i] = v;
        pc += __popc(v);
    }
    pc = warp_sum(pc);
    if ((threadIdx.x & 31) == 0 && pc) atomicAdd(&s_cnt[1], pc);
    __syncthreads();
    if (threadIdx.x == 0) {
        area[m] = s_cnt[0];
        pooled_count[m] = s_cnt[1];
    }
}

// --------------------------------------------------------------------------------------------
// Fused ingest + pooling (W % 32 == 0, at most 2048 patches): the vector pack kernels with a pooling epilogue on the
// words that are still in registers - the packed bits are not read again (pool_packed_kernel re-reads all of them) and
// one launch per chunk disappears.  A block covers 8192 (f32) / 16384 (u8) consecutive pixels of ONE mask, i.e. a few
// image rows: its patches fall into at most a few words of the pooled bitmap, OR-ed into global memory once per block
// (and only by blocks that saw a set pixel).  Same bin rule as pool_packed_kernel, bit-identical results.
// --------------------------------------------------------------------------------------------
constexpr int FUSED_POOL_WORDS = 64;

__device__ __forceinline__ void pool_word(uint32_t word, uint32_t px0, int H, int W, int g, uint32_t* s_pool) {
    // `word` = 32 pixels of one image row starting at flat pixel index px0 (W % 32 == 0); word != 0
    const int y = (int)(px0 / (uint32_t)W);
    const int x0 = (int)(px0 - (uint32_t)y * (uint32_t)W);
    const int c_lo = bin_lo_of32(x0 + __ffs(word) - 1, W, g);   // lowest bin of the first set pixel
    const int c_hi = bin_hi_of32(x0 + 31 - __clz(word), W, g);  // highest bin of the last set pixel
    const int r_lo = bin_lo_of32(y, H, g), r_hi = bin_hi_of32(y, H, g);
    for (int c = c_lo; c <= c_hi; ++c) {
        const int lo = max(bin_start32(c, W, g) - x0, 0), hi = min(bin_end32(c, W, g) - x0, 32);  // window inside the word
        const uint32_t window = (hi - lo) >= 32 ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
        if (word & window)
            for (int r = r_lo; r <= r_hi; ++r) {
                const int b = r * g + c;
                smem_or(&s_pool[b >> 5], 1u << (b & 31));
            }
    }
}

__device__ __forceinline__ void pool_flush(const uint32_t* s_pool, int s_area, int64_t m, int npw,
                                           uint32_t* __restrict__ pooled, int32_t* __restrict__ area) {
    if ((int)threadIdx.x < npw && s_pool[threadIdx.x]) atomicOr(&pooled[m * npw + threadIdx.x], s_pool[threadIdx.x]);
    if (threadIdx.x == 0 && s_area) atomicAdd(&area[m], s_area);
}

__global__ void __launch_bounds__(PACK_THREADS, 6) pack_pool_f32_kernel(const float* __restrict__ masks, int64_t n, int64_t HW,
                                                                      int64_t wpm, uint32_t* __restrict__ bits, int chunks,
                                                                      int H, int W, int g, int npw,
                                                                      uint32_t* __restrict__ pooled,
                                                                      int32_t* __restrict__ area) {
    __shared__ uint32_t s_pool[FUSED_POOL_WORDS];
    __shared__ int s_area;
    const int64_t blk = blockIdx.x;
    const int64_t m = blk / chunks;
    const int chunk = (int)(blk % chunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PX_PER_BLOCK = PACK_THREADS * PACK_UNROLL_F32 * 4;
    const int64_t px_base = (int64_t)chunk * PX_PER_BLOCK + (int64_t)warp * (PACK_UNROLL_F32 * 128);
    const float* src = masks + m * HW;

    uint4 v[PACK_UNROLL_F32];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        const int64_t px = px_base + u * 128 + lane * 4;
        v[u] = (px < HW) ? ldg_stream_u4(src + px) : make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x < FUSED_POOL_WORDS) s_pool[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_area = 0;
    __syncthreads();  // (the loads above are in flight meanwhile)
    uint32_t word[PACK_UNROLL_F32];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        uint32_t nib = (__uint_as_float(v[u].x) > 0.f ? 1u : 0u) | (__uint_as_float(v[u].y) > 0.f ? 2u : 0u) |
                       (__uint_as_float(v[u].z) > 0.f ? 4u : 0u) | (__uint_as_float(v[u].w) > 0.f ? 8u : 0u);
        uint32_t w = nib << (4 * (lane & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        word[u] = w;
    }
    uint32_t out = 0;
#pragma unroll
    for (int u = 0; u < PACK_UNROLL_F32; ++u) {
        uint32_t got = __shfl_sync(0xffffffffu, word[u], (lane & 3) * 8);
        if ((lane >> 2) == u) out = got;
    }
    const int64_t w_idx = px_base / 32 + lane;
    int my_area = 0;
    if (lane < PACK_UNROLL_F32 * 4 && w_idx < wpm) {
        bits[m * wpm + w_idx] = out;
        if (out) {
            my_area = __popc(out);
            pool_word(out, (uint32_t)(w_idx * 32), H, W, g, s_pool);
        }
    }
    my_area = warp_sum(my_area);
    if (lane == 0 && my_area) atomicAdd(&s_area, my_area);
    __syncthreads();
    pool_flush(s_pool, s_area, m, npw, pooled, area);
}

__global__ void __launch_bounds__(PACK_THREADS) pack_pool_u8_kernel(const uint8_t* __restrict__ masks, int64_t n, int64_t HW,
                                                                     int64_t wpm, uint32_t* __restrict__ bits, int chunks,
                                                                     int H, int W, int g, int npw,
                                                                     uint32_t* __restrict__ pooled,
                                                                     int32_t* __restrict__ area) {
    __shared__ uint32_t s_pool[FUSED_POOL_WORDS];
    __shared__ int s_area;
    const int64_t blk = blockIdx.x;
    const int64_t m = blk / chunks;
    const int chunk = (int)(blk % chunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PX_PER_BLOCK = PACK_THREADS * PACK_UNROLL * 16;
    const int64_t px_base = (int64_t)chunk * PX_PER_BLOCK + (int64_t)warp * (PACK_UNROLL * 512);
    const uint8_t* src = masks + m * HW;

    uint4 v[PACK_UNROLL];
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
        const int64_t px = px_base + u * 512 + lane * 16;
        v[u] = (px < HW) ? ldg_stream_u4(src + px) : make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x < FUSED_POOL_WORDS) s_pool[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_area = 0;
    __syncthreads();
    int my_area = 0;
#pragma unroll
    for (int u = 0; u < PACK_UNROLL; ++u) {
        auto nib = [](uint32_t x) -> uint32_t {
            uint32_t mbits = __vcmpgtu4(x, 0u) & 0x01010101u;
            return (mbits * 0x01020408u) >> 24 & 0xfu;
        };
        uint32_t half = nib(v[u].x) | (nib(v[u].y) << 4) | (nib(v[u].z) << 8) | (nib(v[u].w) << 12);
        uint32_t w = half << (16 * (lane & 1));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        const int64_t w_idx = px_base / 32 + u * 16 + (lane >> 1);
        if ((lane & 1) == 0 && w_idx < wpm) {
            bits[m * wpm + w_idx] = w;
            if (w) {
                my_area += __popc(w);
                pool_word(w, (uint32_t)(w_idx * 32), H, W, g, s_pool);
            }
        }
    }
    my_area = warp_sum(my_area);
    if (lane == 0 && my_area) atomicAdd(&s_area, my_area);
    __syncthreads();
    pool_flush(s_pool, s_area, m, npw, pooled, area);
}

// pooled patch count of every mask: one warp per mask
__global__ void __launch_bounds__(256) pooled_count_kernel(const uint32_t* __restrict__ pooled, int64_t n, int npw,
                                                            int32_t* __restrict__ pooled_count) {
    const int lane = threadIdx.x & 31;
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= n) return;
    int c = 0;
    for (int w = lane; w < npw; w += 32) c += __popc(pooled[m * npw + w]);
    c = warp_sum(c);
    if (lane == 0) pooled_count[m] = c;
}

// --------------------------------------------------------------------------------------------
// region sums: one warp per proposal; union count: one block per episode.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void region_sums_block(const uint32_t* __restrict__ pooled, int64_t total, int P, int N, int npw,
                                                  const float* __restrict__ vva, const float* __restrict__ vta,
                                                  float* __restrict__ sum_vva, float* __restrict__ sum_vta) {
    const int lane = threadIdx.x & 31;
    const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wg >= total) return;
    const int64_t e = wg / P;
    const uint32_t* row = pooled + wg * npw;
    const float* a = vva + e * N;
    const float* t = vta + e * N;
    double sa = 0.0, st = 0.0;
    for (int w = 0; w < npw; ++w) {
        const uint32_t word = row[w];
        const int b = w * 32 + lane;
        if (((word >> lane) & 1u) && b < N) {
            sa += (double)a[b];
            st += (double)t[b];
        }
    }
    sa = warp_sum(sa);
    st = warp_sum(st);
    if (lane == 0) {
        sum_vva[wg] = (float)sa;
        sum_vta[wg] = (float)st;
    }
}

// 256 threads per episode: thread t ORs word (t % 64 ...) over a slice of the proposals, then a
// shared-memory OR per word and a popcount.
__device__ __forceinline__ void union_count_block(const uint32_t* __restrict__ pooled, int64_t e, int P, int npw,
                                                  int32_t* __restrict__ union_count) {
    extern __shared__ uint32_t s_union[];  // npw words + 1 counter
    for (int i = threadIdx.x; i <= npw; i += blockDim.x) s_union[i] = 0;
    __syncthreads();
    const uint32_t* base = pooled + e * P * npw;
    const int total = P * npw;
    // eight independent loads in flight per thread: one CTA per episode is latency-bound otherwise
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            v[u] = i < total ? __ldg(base + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (v[u]) smem_or(&s_union[(i0 + u * blockDim.x) % npw], v[u]);
    }
    __syncthreads();
    int c = 0;
    for (int w = threadIdx.x; w < npw; w += blockDim.x) c += __popc(s_union[w]);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<int*>(&s_union[npw]), c);
    __syncthreads();
    if (threadIdx.x == 0) union_count[e] = (int)s_union[npw];
}

// One launch for both: the first `region_blocks` blocks compute the per-proposal region sums (8 proposals each), the
// remaining E blocks the union patch count of one episode each.
__global__ void __launch_bounds__(256) region_union_kernel(const uint32_t* __restrict__ pooled, int64_t total, int P, int N,
                                                           int npw, const float* __restrict__ vva,
                                                           const float* __restrict__ vta, float* __restrict__ sum_vva,
                                                           float* __restrict__ sum_vta, int32_t* __restrict__ union_count,
                                                           unsigned region_blocks) {
    if (blockIdx.x < region_blocks) region_sums_block(pooled, total, P, N, npw, vva, vta, sum_vva, sum_vta);
    else union_count_block(pooled, (int64_t)(blockIdx.x - region_blocks), P, npw, union_count);
}

// --------------------------------------------------------------------------------------------
// pairwise intersections, AND + popcount.  A block owns a 64x64 tile of pairs over a slice of the
// words; operand slices are staged k-major in shared memory (padded: conflict-free both ways);
// each thread keeps a 4x4 tile of int32 counters.  Integer atomics make the split-K reduction exact
// and order independent.
// --------------------------------------------------------------------------------------------
constexpr int PW_TILE = 64;
constexpr int PW_KC = 32;

__global__ void __launch_bounds__(256) pairwise_popc_kernel(const uint32_t* __restrict__ bits, int P, int64_t wpm,
                                                            int tiles, int ksplit, int64_t words_per_split,
                                                            int32_t* __restrict__ inter) {
    __shared__ uint32_t sA[PW_KC][PW_TILE + 1];
    __shared__ uint32_t sB[PW_KC][PW_TILE + 1];
    // blockIdx.x -> (upper-triangular tile pair), blockIdx.y -> k split, blockIdx.z -> episode
    int t = blockIdx.x, ti = 0;
    while (t >= tiles - ti) {
        t -= tiles - ti;
        ++ti;
    }
    const int tj = ti + t;
    const int64_t e = blockIdx.z;
    const int64_t k_begin = (int64_t)blockIdx.y * words_per_split;
    const int64_t k_end = min(k_begin + words_per_split, wpm);
    const uint32_t* base = bits + e * P * wpm;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    int acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0;

    const int lk = threadIdx.x & 31;  // word within the chunk
    const int lr = threadIdx.x >> 5;  // 8 rows per pass
    for (int64_t k0 = k_begin; k0 < k_end; k0 += PW_KC) {
#pragma unroll
        for (int r = 0; r < PW_TILE; r += 8) {
            const int i = ti * PW_TILE + r + lr, j = tj * PW_TILE + r + lr;
            const int64_t k = k0 + lk;
            sA[lk][r + lr] = (i < P && k < k_end) ? base[(int64_t)i * wpm + k] : 0u;
            sB[lk][r + lr] = (j < P && k < k_end) ? base[(int64_t)j * wpm + k] : 0u;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < PW_KC; ++k) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = sA[k][ty + 16 * q];
                b[q] = sB[k][tx + 16 * q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] += __popc(a[p] & b[q]);
        }
        __syncthreads();
    }
    int32_t* out = inter + e * P * (int64_t)P;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ti * PW_TILE + ty + 16 * p, j = tj * PW_TILE + tx + 16 * q;
            if (i < P && j < P && acc[p][q] != 0) {
                if (ti == tj) {
                    atomicAdd(&out[(int64_t)i * P + j], acc[p][q]);
                } else {
                    atomicAdd(&out[(int64_t)i * P + j], acc[p][q]);
                    atomicAdd(&out[(int64_t)j * P + i], acc[p][q]);
                }
            }
        }
}

// --------------------------------------------------------------------------------------------
// merge: OR of the selected rows, optional expansion to float32 0/1.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_masks_kernel(const uint32_t* __restrict__ bits,
                                                          const uint8_t* __restrict__ flags, int P, int64_t wpm,
                                                          int64_t HW, uint32_t* __restrict__ merged_bits,
                                                          float* __restrict__ merged_f32) {
    extern __shared__ int s_sel[];  // P indices + 1 counter
    int* s_n = s_sel + P;
    const int64_t e = blockIdx.y;
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x)
        if (flags[e * P + p] & 2) s_sel[atomicAdd(s_n, 1)] = p;
    __syncthreads();
    const int nsel = *s_n;
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    if (w < wpm) {
        const uint32_t* base = bits + e * P * wpm + w;
        for (int k = 0; k < nsel; ++k) acc |= base[(int64_t)s_sel[k] * wpm];
        if (merged_bits) merged_bits[e * wpm + w] = acc;
    }
    if (merged_f32) {
        const int64_t w_warp = w - lane;  // first word of this warp
        float* out = merged_f32 + e * HW;
#pragma unroll 4
        for (int l = 0; l < 32; ++l) {
            const uint32_t wv = __shfl_sync(0xffffffffu, acc, l);
            const int64_t px = (w_warp + l) * 32 + lane;
            if (px < HW) out[px] = (float)((wv >> lane) & 1u);
        }
    }
}

// --------------------------------------------------------------------------------------------
// Matcher: matched points inside each packed mask (one block per mask).
// --------------------------------------------------------------------------------------------
__global__ void points_in_masks_kernel(const uint32_t* __restrict__ bits, int H, int W, int64_t wpm,
                                       const int32_t* __restrict__ points, int K, int32_t* __restrict__ out) {
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    const int64_t m = blockIdx.x;
    int c = 0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int x = min(max(points[2 * k], 0), W - 1);
        const int y = min(max(points[2 * k + 1], 0), H - 1);
        const int64_t px = (int64_t)y * W + x;
        c += (bits[m * wpm + (px >> 5)] >> (px & 31)) & 1u;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_total, c);
    __syncthreads();
    if (threadIdx.x == 0) out[m] = s_total;
}

__global__ void matcher_scores_kernel(const int32_t* __restrict__ points_in, const int32_t* __restrict__ pooled_count,
                                      const float* __restrict__ emd, int64_t n, int K, float alpha, float beta,
                                      float expo, float* __restrict__ purity, float* __restrict__ coverage,
                                      float* __restrict__ scores) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // python float (double) quotient -> float32 tensor -> + 1e-6 in float32 (Matcher.py:1203-1207)
    const double area = fmax((double)pooled_count[i], 1.0);
    const float pur = (float)((double)points_in[i] / area) + 1e-6f;
    const float cov = (float)((double)points_in[i] / (double)K) + 1e-6f;
    purity[i] = pur;
    coverage[i] = cov;
    scores[i] = alpha * emd[i] + beta * pur * powf(cov, expo);
}

// --------------------------------------------------------------------------------------------
// evaluator areas: histc(bins=2, min=0, max=1) semantics (values outside [0,1] are not counted).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eval_areas_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                         const float* __restrict__ ignore, int64_t HW,
                                                         int32_t* __restrict__ out) {
    __shared__ int s_c[6];
    if (threadIdx.x < 6) s_c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t m = blockIdx.y;
    int c[6] = {0, 0, 0, 0, 0, 0};  // inter_bg, inter_fg, pred_bg, pred_fg, gt_bg, gt_fg
    for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < HW; px += (int64_t)gridDim.x * blockDim.x) {
        float g = gt[m * HW + px];
        float p = pred[m * HW + px];
        if (ignore) {
            g = g + ignore[m * HW + px] * 255.f;
            if (g == 255.f) p = 255.f;
        }
        auto bin = [](float v) -> int { return (v >= 0.f && v <= 1.f) ? (v >= 0.5f ? 1 : 0) : -1; };
        const int bp = bin(p), bg = bin(g);
        if (p == g && bp >= 0) c[bp]++;
        if (bp >= 0) c[2 + bp]++;
        if (bg >= 0) c[4 + bg]++;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int v = warp_sum(c[i]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[i], v);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int inter = s_c[threadIdx.x];
        atomicAdd(&out[m * 4 + threadIdx.x], inter);
        atomicAdd(&out[m * 4 + 2 + threadIdx.x], s_c[2 + threadIdx.x] + s_c[4 + threadIdx.x] - inter);
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int64_t marsb200_words_per_mask(int64_t hw) { return ceil_div64(ceil_div64(hw, 32), 32) * 32; }

// The ingest kernels (no shared memory) and pool_packed_kernel (~8 KB per CTA) ask for the SAME L1 / shared-memory split:
// CTAs of two kernels only share an SM when their carveouts agree, and the episode engine runs the
// issue-bound pooling of one chunk beside the load-path-bound ingest of the next on one SM partition.
static cudaError_t set_ingest_carveouts() {
    const int pct = 50;
    cudaError_t e = cudaFuncSetAttribute(pack_f32_vec_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pack_u8_vec_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pool_packed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    return e;
}
static PerDeviceOnce g_ingest_carveouts;

int marsb200_pack_masks(const void* masks, int mask_dtype, int64_t n, int64_t HW, uint32_t* bits, void* stream) {
    MARS_REQUIRE(masks && bits, "null pointer");
    MARS_CUDA_OK(per_device_once(g_ingest_carveouts, set_ingest_carveouts));
    MARS_REQUIRE(n > 0 && HW > 0, "empty input");
    MARS_REQUIRE(mask_dtype == MARSB200_MASK_F32 || mask_dtype == MARSB200_MASK_U8, "mask_dtype");
    const int64_t wpm = marsb200_words_per_mask(HW);
    cudaStream_t s = as_stream(stream);
    const bool aligned = (reinterpret_cast<uintptr_t>(masks) & 15) == 0;
    if (mask_dtype == MARSB200_MASK_F32 && aligned && HW % 4 == 0) {
        const int chunks = (int)ceil_div64(wpm * 32, PACK_THREADS * PACK_UNROLL_F32 * 4);
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_f32_vec_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const float*)masks, n, HW, wpm, bits, chunks);
    } else if (mask_dtype == MARSB200_MASK_U8 && aligned && HW % 16 == 0) {
        const int chunks = (int)ceil_div64(wpm * 32, PACK_THREADS * PACK_UNROLL * 16);
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_u8_vec_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const uint8_t*)masks, n, HW, wpm, bits, chunks);
    } else {
        const int64_t warps = n * wpm;
        const unsigned grid = (unsigned)std::min<int64_t>(ceil_div64(warps, 8), 148 * 64);
        if (mask_dtype == MARSB200_MASK_F32)
            pack_scalar_kernel<float><<<grid, 256, 0, s>>>((const float*)masks, n, HW, wpm, bits);
        else
            pack_scalar_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)masks, n, HW, wpm, bits);
    }
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

// Words [word_begin, word_begin + word_count) of every mask: the vector kernels on shifted base pointers (the row strides HW
// and wpm are unchanged; a slice of whole 512-word blocks inside the mask's pixels never meets the tail handling).
int marsb200_pack_masks_slice(const void* masks, int mask_dtype, int64_t n, int64_t HW, int64_t word_begin, int64_t word_count,
                              uint32_t* bits, void* stream) {
    MARS_REQUIRE(masks && bits, "null pointer");
    MARS_CUDA_OK(per_device_once(g_ingest_carveouts, set_ingest_carveouts));
    MARS_REQUIRE(n > 0 && HW > 0, "empty input");
    MARS_REQUIRE(mask_dtype == MARSB200_MASK_F32 || mask_dtype == MARSB200_MASK_U8, "mask_dtype");
    const int64_t wpm = marsb200_words_per_mask(HW);
    MARS_REQUIRE(word_begin >= 0 && word_count > 0 && word_begin % 512 == 0 && word_count % 512 == 0,
                 "slice must be whole 512-word blocks");
    MARS_REQUIRE((word_begin + word_count) * 32 <= HW, "slice must lie inside the mask's pixels");
    const int64_t esz = mask_dtype == MARSB200_MASK_F32 ? 4 : 1;
    const char* src = static_cast<const char*>(masks) + word_begin * 32 * esz;
    MARS_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && HW % (mask_dtype == MARSB200_MASK_F32 ? 4 : 16) == 0,
                 "slices need 16-byte aligned masks with HW % 4 == 0 (float32) / HW % 16 == 0 (uint8)");
    cudaStream_t s = as_stream(stream);
    if (mask_dtype == MARSB200_MASK_F32) {
        const int chunks = (int)(word_count * 32 / (PACK_THREADS * PACK_UNROLL_F32 * 4));
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_f32_vec_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const float*)src, n, HW, wpm, bits + word_begin, chunks);
    } else {
        const int chunks = (int)(word_count * 32 / (PACK_THREADS * PACK_UNROLL * 16));
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_u8_vec_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const uint8_t*)src, n, HW, wpm, bits + word_begin, chunks);
    }
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_pool_mask(const void* masks, int mask_dtype, int64_t n, int H, int W, int g, uint8_t* out, void* stream) {
    MARS_REQUIRE(masks && out, "null pointer");
    MARS_REQUIRE(n > 0 && H > 0 && W > 0 && g > 0 && g <= H && g <= W, "shape");
    MARS_REQUIRE(g <= 255 && n * g < (1ll << 31), "g <= 255");
    const unsigned grid = (unsigned)(n * g);
    if (mask_dtype == MARSB200_MASK_F32)
        pool_mask_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)masks, n, H, W, g, out);
    else if (mask_dtype == MARSB200_MASK_U8)
        pool_mask_kernel<uint8_t><<<grid, 256, 0, as_stream(stream)>>>((const uint8_t*)masks, n, H, W, g, out);
    else
        MARS_REQUIRE(false, "mask_dtype");
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_pool_packed(const uint32_t* bits, int64_t n, int H, int W, int g, uint32_t* pooled, int32_t* area,
                         int32_t* pooled_count, void* stream) {
    MARS_REQUIRE(bits && pooled && area && pooled_count, "null pointer");
    MARS_REQUIRE(n > 0 && H > 0 && W > 0 && g > 0 && g <= H && g <= W, "shape");
    const int npw = ceil_div(g * g, 32);
    const int64_t wpm = marsb200_words_per_mask((int64_t)H * W);
    MARS_REQUIRE(n < (1ll << 31), "too many masks");
    MARS_REQUIRE((int64_t)H * W < (1ll << 31), "mask too large");
    int rw = W / 32 + 2;  // row-aligned bit row: ceil(W/32) words + 1 spill word
    if (W % 128 == 0) rw = (rw + 3) / 4 * 4;  // 16-byte aligned bit rows for the 128-bit fast path
    MARS_REQUIRE(g <= 255 && H <= 32768 && W <= 32768, "g <= 255, H, W <= 32768");
    const size_t smem = ((size_t)g * rw + npw + 2) * sizeof(uint32_t) + (size_t)H * sizeof(uint16_t);
    MARS_REQUIRE(smem <= 48 * 1024, "g * W too large for the pooling scratch");
    MARS_CUDA_OK(per_device_once(g_ingest_carveouts, set_ingest_carveouts));
    pool_packed_kernel<<<(unsigned)n, POOL_THREADS, smem, as_stream(stream)>>>(bits, n, H, W, g, wpm, npw, rw, pooled,
                                                                             area, pooled_count);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_pack_pool_masks(const void* masks, int mask_dtype, int64_t n, int H, int W, int g, uint32_t* bits,
                             uint32_t* pooled, int32_t* area, int32_t* pooled_count, void* stream) {
    MARS_REQUIRE(masks && bits && pooled && area && pooled_count, "null pointer");
    MARS_REQUIRE(n > 0 && H > 0 && W > 0 && g > 0 && g <= H && g <= W, "shape");
    MARS_REQUIRE(mask_dtype == MARSB200_MASK_F32 || mask_dtype == MARSB200_MASK_U8, "mask_dtype");
    const int64_t HW = (int64_t)H * W;
    const int64_t wpm = marsb200_words_per_mask(HW);
    const int npw = ceil_div(g * g, 32);
    const bool aligned = (reinterpret_cast<uintptr_t>(masks) & 15) == 0;
    const bool vec = aligned && (mask_dtype == MARSB200_MASK_F32 ? HW % 4 == 0 : HW % 16 == 0);
    const bool fusable = vec && W % 32 == 0 && npw <= FUSED_POOL_WORDS && H <= 32768 && W <= 32768 && g <= 255 &&
                         HW < (1ll << 31);
    if (!fusable) {  // any other geometry: the two kernels (same results)
        if (int rc = marsb200_pack_masks(masks, mask_dtype, n, HW, bits, stream)) return rc;
        return marsb200_pool_packed(bits, n, H, W, g, pooled, area, pooled_count, stream);
    }
    cudaStream_t s = as_stream(stream);
    MARS_CUDA_OK(cudaMemsetAsync(pooled, 0, sizeof(uint32_t) * (size_t)n * npw, s));
    MARS_CUDA_OK(cudaMemsetAsync(area, 0, sizeof(int32_t) * (size_t)n, s));
    if (mask_dtype == MARSB200_MASK_F32) {
        const int chunks = (int)ceil_div64(wpm * 32, PACK_THREADS * PACK_UNROLL_F32 * 4);
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_pool_f32_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const float*)masks, n, HW, wpm, bits, chunks, H,
                                                                           W, g, npw, pooled, area);
    } else {
        const int chunks = (int)ceil_div64(wpm * 32, PACK_THREADS * PACK_UNROLL * 16);
        MARS_REQUIRE(n * chunks < (1ll << 31), "grid too large");
        pack_pool_u8_kernel<<<(unsigned)(n * chunks), PACK_THREADS, 0, s>>>((const uint8_t*)masks, n, HW, wpm, bits, chunks, H,
                                                                          W, g, npw, pooled, area);
    }
    MARS_LAUNCH_OK();
    pooled_count_kernel<<<(unsigned)ceil_div64(n * 32, 256), 256, 0, s>>>(pooled, n, npw, pooled_count);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_region_sums(const uint32_t* pooled, int E, int P, int N, const float* vva, const float* vta,
                         float* sum_vva, float* sum_vta, int32_t* union_count, void* stream) {
    MARS_REQUIRE(pooled && vva && vta && sum_vva && sum_vta && union_count, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && N > 0, "shape");
    const int npw = ceil_div(N, 32);
    const int64_t total = (int64_t)E * P;
    const unsigned region_blocks = (unsigned)ceil_div64(total, 8);
    region_union_kernel<<<region_blocks + (unsigned)E, 256, (npw + 1) * sizeof(uint32_t), as_stream(stream)>>>(
        pooled, total, P, N, npw, vva, vta, sum_vva, sum_vta, union_count, region_blocks);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

namespace marsb200 {
int pairwise_popc(const uint32_t* bits, int E, int P, int64_t wpm, int32_t* inter, cudaStream_t s) {
    const int tiles = ceil_div(P, PW_TILE);
    const int tile_pairs = tiles * (tiles + 1) / 2;
    // enough blocks for ~4 waves of 148 SMs x 4 resident blocks, but at least 4 chunks of work per block
    int ksplit = (int)std::min<int64_t>(std::max<int64_t>(1, (148 * 16) / ((int64_t)tile_pairs * E)), std::max<int64_t>(1, wpm / (4 * PW_KC)));
    ksplit = min(ksplit, 65535);
    int64_t words_per_split = ceil_div64(ceil_div64(wpm, ksplit), PW_KC) * PW_KC;
    ksplit = (int)ceil_div64(wpm, words_per_split);
    MARS_CUDA_OK(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)E * P * P, s));
    dim3 grid(tile_pairs, ksplit, E);
    pairwise_popc_kernel<<<grid, 256, 0, s>>>(bits, P, wpm, tiles, ksplit, words_per_split, inter);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}
}  // namespace marsb200

extern "C" {

int marsb200_merge_masks(const uint32_t* bits, const uint8_t* flags, int E, int P, int64_t HW, uint32_t* merged_bits,
                         float* merged_f32, void* stream) {
    MARS_REQUIRE(bits && flags, "null pointer");
    MARS_REQUIRE(merged_bits || merged_f32, "no output requested");
    MARS_REQUIRE(E > 0 && P > 0 && HW > 0 && E <= 65535, "shape");
    const int64_t wpm = marsb200_words_per_mask(HW);
    dim3 grid((unsigned)ceil_div64(wpm, 256), E);
    merge_masks_kernel<<<grid, 256, (P + 1) * sizeof(int), as_stream(stream)>>>(bits, flags, P, wpm, HW, merged_bits,
                                                                                merged_f32);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_points_in_masks(const uint32_t* bits, int64_t n, int H, int W, const int32_t* points, int K, int32_t* out,
                             void* stream) {
    MARS_REQUIRE(bits && points && out, "null pointer");
    MARS_REQUIRE(n > 0 && K > 0 && H > 0 && W > 0, "shape");
    const int64_t wpm = marsb200_words_per_mask((int64_t)H * W);
    points_in_masks_kernel<<<(unsigned)n, 128, 0, as_stream(stream)>>>(bits, H, W, wpm, points, K, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_matcher_scores(const int32_t* points_in, const int32_t* pooled_count, const float* emd, int64_t n, int K,
                            float alpha, float beta, float expo, float* purity, float* coverage, float* scores,
                            void* stream) {
    MARS_REQUIRE(points_in && pooled_count && emd && purity && coverage && scores, "null pointer");
    MARS_REQUIRE(n > 0 && K > 0, "shape");
    matcher_scores_kernel<<<(unsigned)ceil_div64(n, 128), 128, 0, as_stream(stream)>>>(
        points_in, pooled_count, emd, n, K, alpha, beta, expo, purity, coverage, scores);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_eval_areas(const float* pred, const float* gt, const float* ignore, int64_t n, int64_t HW, int32_t* out,
                        void* stream) {
    MARS_REQUIRE(pred && gt && out, "null pointer");
    MARS_REQUIRE(n > 0 && n <= 65535 && HW > 0, "shape");
    MARS_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int32_t) * 4 * (size_t)n, as_stream(stream)));
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div64(HW, 256 * 8), 148 * 4);
    eval_areas_kernel<<<dim3(gx, (unsigned)n), 256, 0, as_stream(stream)>>>(pred, gt, ignore, HW, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

// Proposal wire format and SAM automatic-mask-generator post-processing: the step in front of
// the ranking path.  Uncompressed COCO RLE (column-major runs, segment_anything/utils/amg.py:107-148) decoded
// straight into the packed row-major bit format every mask kernel consumes, boxes from packed masks
// (amg.py:310-353), the stability score (amg.py:156-176) and greedy box NMS with torchvision semantics
// (segment_anything/automatic_mask_generator.py:370-376).
#include <algorithm>

#include "common.cuh"

namespace marsb200 {

// --------------------------------------------------------------------------------------------
// RLE -> packed bits, two passes over packed data only:
//   (1) every run of ones [s, e) of the column-major pixel order becomes a range of bits in a column-major bit
//       matrix colbits[m][x][H/32] (whole words stored, edge words OR-ed);
//   (2) 32 x 32 bit tiles are transposed with ballots into the row-major layout bits[m][y][W/32].
// One CTA per mask in (1): the run starts are a block-wide running prefix sum of the counts.
// --------------------------------------------------------------------------------------------
constexpr int RLE_THREADS = 256;

__global__ void __launch_bounds__(RLE_THREADS) rle_fill_kernel(const int32_t* __restrict__ counts,
                                                               const int64_t* __restrict__ offsets, int64_t HW,
                                                               int64_t col_words, uint32_t* __restrict__ colbits,
                                                               int* __restrict__ status) {
    __shared__ long long s_warp[RLE_THREADS / 32];
    __shared__ long long s_base;
    const int64_t m = blockIdx.x;
    const int32_t* c = counts + offsets[m];
    const int64_t n_runs = offsets[m + 1] - offsets[m];
    uint32_t* out = colbits + m * col_words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int64_t r0 = 0; r0 < n_runs; r0 += RLE_THREADS) {
        const int64_t r = r0 + tid;
        const long long len = r < n_runs ? (long long)c[r] : 0;
        // inclusive scan of the chunk
        long long incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        long long off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        const long long start = off + incl - len, end = off + incl;
        if (r < n_runs && (len < 0 || end > HW)) atomicExch(status, 1);  // malformed RLE
        if (r < n_runs && (r & 1) && len > 0 && end <= HW) {  // odd runs are ones (counts start with a zero run)
            const long long w0 = start >> 5, w1 = (end - 1) >> 5;
            const uint32_t first = 0xffffffffu << (start & 31);
            const uint32_t last = 0xffffffffu >> (31 - ((end - 1) & 31));
            if (w0 == w1) {
                atomicOr(out + w0, first & last);
            } else {
                atomicOr(out + w0, first);
                for (long long w = w0 + 1; w < w1; ++w) out[w] = 0xffffffffu;
                atomicOr(out + w1, last);
            }
        }
        __syncthreads();
        if (tid == RLE_THREADS - 1) s_base = end;
        __syncthreads();
    }
    if (tid == 0 && s_base != HW) atomicExch(status, 1);  // the counts must cover the image exactly
}

// colbits [n][W][H/32] -> bits [n][H][W/32] (+ zero padding words up to wpm).  One CTA transposes a block of 8 x 8 tiles of
// 32 x 32 bits (256 columns x 256 rows): thread t reads the 8 consecutive words of column x0 + t (one 32-byte sector), the
// tiles are transposed with warp shuffles out of shared memory, thread t writes the 8 consecutive words of row y0 + t.  Every
// global access is a full sector (the one-warp-per-tile version read and wrote 4 useful bytes per sector).
constexpr int BT_TILES = 8;

__global__ void __launch_bounds__(256) bit_transpose_kernel(const uint32_t* __restrict__ colbits, int H, int W, int64_t wpm,
                                                            uint32_t* __restrict__ bits) {
    __shared__ uint32_t s_in[32 * BT_TILES][BT_TILES + 1];   // [column][tile row], padded against bank conflicts
    __shared__ uint32_t s_out[32 * BT_TILES][BT_TILES + 1];  // [row][tile column]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw32 = H >> 5, ww32 = W >> 5;
    const int bx = blockIdx.x % ((ww32 + BT_TILES - 1) / BT_TILES), by = blockIdx.x / ((ww32 + BT_TILES - 1) / BT_TILES);
    const int64_t m = blockIdx.y;
    const int tx0 = bx * BT_TILES, ty0 = by * BT_TILES;
    {   // column x0 + tid: words ty0 .. ty0 + 7
        const int x = tx0 * 32 + tid;
        const uint32_t* src = colbits + (m * W + x) * hw32 + ty0;
        if (x < W && (hw32 & 3) == 0 && ty0 + BT_TILES <= hw32) {  // 16-byte aligned: two 128-bit loads
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(src)), b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
            s_in[tid][0] = a.x; s_in[tid][1] = a.y; s_in[tid][2] = a.z; s_in[tid][3] = a.w;
            s_in[tid][4] = b.x; s_in[tid][5] = b.y; s_in[tid][6] = b.z; s_in[tid][7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < BT_TILES; ++k) s_in[tid][k] = (x < W && ty0 + k < hw32) ? src[k] : 0u;
        }
    }
    __syncthreads();
    // warp w owns tile row ty0 + w; for each tile column the 32 lanes (= 32 columns) transpose their word with ballots
#pragma unroll 2
    for (int k = 0; k < BT_TILES; ++k) {
        // 32 x 32 bit transpose across the warp in five exchange steps (block swaps of 16, 8, 4, 2, 1): lane l enters with
        // column x0 + l (bit b = row y0 + b) and leaves with row y0 + l (bit b = column x0 + b).  ~30 instructions per tile;
        // the 32-ballot version this replaces was issue-bound at 0.8 TB/s.
        uint32_t a = s_in[k * 32 + lane][warp];
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
            const uint32_t o = __shfl_xor_sync(0xffffffffu, a, j);
            if ((lane & j) == 0) a ^= (((a >> j) ^ o) & m) << j;
            else a ^= ((o >> j) ^ a) & m;
        }
        s_out[warp * 32 + lane][k] = a;
    }
    __syncthreads();
    {   // row y0 + tid: words tx0 .. tx0 + 7
        const int y = ty0 * 32 + tid;
        if (y < H) {
            uint32_t* dst = bits + m * wpm + (int64_t)y * ww32 + tx0;
            if ((ww32 & 3) == 0 && (wpm & 3) == 0 && tx0 + BT_TILES <= ww32) {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(s_out[tid][0], s_out[tid][1], s_out[tid][2], s_out[tid][3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(s_out[tid][4], s_out[tid][5], s_out[tid][6], s_out[tid][7]);
            } else {
#pragma unroll
                for (int k = 0; k < BT_TILES; ++k)
                    if (tx0 + k < ww32) dst[k] = s_out[tid][k];
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// XYXY boxes of packed masks, [0, 0, 0, 0] for an empty mask (batched_mask_to_box).  One CTA per mask;
// needs W % 32 == 0 so that rows are word aligned (the generic path walks bit by bit per row).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_boxes_kernel(const uint32_t* __restrict__ bits, int64_t wpm, int H, int W,
                                                          int32_t* __restrict__ boxes) {
    __shared__ int s_top, s_bottom, s_left, s_right;
    const int64_t m = blockIdx.x;
    const uint32_t* b = bits + m * wpm;
    if (threadIdx.x == 0) {
        s_top = H;
        s_bottom = -1;
        s_left = W;
        s_right = -1;
    }
    __syncthreads();
    int top = H, bottom = -1, left = W, right = -1;
    const int64_t hw = (int64_t)H * W;
    const int64_t words = (hw + 31) >> 5;
    for (int64_t w = threadIdx.x; w < words; w += blockDim.x) {
        uint32_t v = b[w];
        while (v) {
            const int bit = __ffs(v) - 1;
            v &= v - 1;
            const int64_t px = (w << 5) + bit;
            const int y = (int)(px / W), x = (int)(px - (int64_t)y * W);
            top = min(top, y);
            bottom = max(bottom, y);
            left = min(left, x);
            right = max(right, x);
            // within one word of a word-aligned row only the extreme bits matter: skip to the highest set bit
            if ((W & 31) == 0 && v) {
                const int hb = 31 - __clz(v);
                right = max(right, (int)(((w << 5) + hb) - (int64_t)y * W));
                v = 0;
            }
        }
    }
    top = -warp_max((float)-top);
    bottom = (int)warp_max((float)bottom);
    left = -(int)warp_max((float)-left);
    right = (int)warp_max((float)right);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_top, top);
        atomicMax(&s_bottom, bottom);
        atomicMin(&s_left, left);
        atomicMax(&s_right, right);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool empty = s_right < s_left || s_bottom < s_top;
        boxes[4 * m + 0] = empty ? 0 : s_left;
        boxes[4 * m + 1] = empty ? 0 : s_top;
        boxes[4 * m + 2] = empty ? 0 : s_right;
        boxes[4 * m + 3] = empty ? 0 : s_bottom;
    }
}

// --------------------------------------------------------------------------------------------
// Stability score: |logits > thr + off| / |logits > thr - off| (calculate_stability_score); integer counts,
// one float32 division.  grid (chunks, n).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stability_counts_kernel(const float* __restrict__ logits, int64_t HW, float hi,
                                                                float lo, int32_t* __restrict__ counts) {
    const int64_t m = blockIdx.y;
    const float* src = logits + m * HW;
    int a = 0, b = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int64_t nv = HW >> 2;
        for (; i < nv; i += stride) {
            const uint4 u = ldg_stream_u4(src + 4 * i);
            const float v[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a += v[k] > hi ? 1 : 0;
                b += v[k] > lo ? 1 : 0;
            }
        }
    } else {
        for (; i < HW; i += stride) {
            const float v = src[i];
            a += v > hi ? 1 : 0;
            b += v > lo ? 1 : 0;
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counts + 2 * m, a);
        atomicAdd(counts + 2 * m + 1, b);
    }
}

__global__ void stability_finish_kernel(const int32_t* __restrict__ counts, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fdiv_rn((float)counts[2 * i], (float)counts[2 * i + 1]);  // 0/0 -> NaN like torch
}

// --------------------------------------------------------------------------------------------
// Greedy box NMS (torchvision.ops.nms semantics: boxes sorted by score descending; a box is dropped iff its IoU
// with an already kept higher-scored box is > threshold; IoU = inter / (area_a + area_b - inter) in float32).
// One CTA: stable bitonic rank (score desc, index asc), suppression bit matrix in global workspace, warp scan.
// --------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 1024;

__device__ __forceinline__ bool box_iou_gt(const float4 a, const float4 b, float thr) {
    const float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x), ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    const float inter = fmaxf(iw, 0.f) * fmaxf(ih, 0.f);
    const float area_a = (a.z - a.x) * (a.w - a.y), area_b = (b.z - b.x) * (b.w - b.y);
    return inter / (area_a + area_b - inter) > thr;
}

__global__ void __launch_bounds__(NMS_THREADS) box_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                               int n, int n_pow2, float thr, int32_t* __restrict__ order,
                                                               unsigned long long* __restrict__ supp, uint8_t* __restrict__ keep,
                                                               int32_t* __restrict__ n_keep) {
    extern __shared__ unsigned char nms_smem[];
    float* s_key = reinterpret_cast<float*>(nms_smem);         // [n_pow2]
    int* s_idx = reinterpret_cast<int*>(s_key + n_pow2);        // [n_pow2]
    const int tid = threadIdx.x;
    for (int i = tid; i < n_pow2; i += NMS_THREADS) {
        s_key[i] = i < n ? scores[i] : -INFINITY;
        s_idx[i] = i < n ? i : 0x7fffffff;
    }
    __syncthreads();
    // bitonic sort: descending score, ascending index among equal scores (NaN scores are not supported)
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pow2; i += NMS_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const float ka = s_key[i], kb = s_key[p];
                    const int ia = s_idx[i], ib = s_idx[p];
                    const bool a_first = ka > kb || (ka == kb && ia < ib);
                    const bool up = (i & k) == 0;
                    if (up ? !a_first : a_first) {
                        s_key[i] = kb;
                        s_key[p] = ka;
                        s_idx[i] = ib;
                        s_idx[p] = ia;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += NMS_THREADS) order[i] = s_idx[i];
    // suppression matrix in rank space: bit c of supp[r][w] set iff rank 64w+c > r overlaps rank r above the threshold
    const int words = (n + 63) >> 6;
    const float4* b4 = reinterpret_cast<const float4*>(boxes);
    for (int64_t t = tid; t < (int64_t)n * words; t += NMS_THREADS) {
        const int r = (int)(t / words), w = (int)(t % words);
        const float4 a = b4[s_idx[r]];
        unsigned long long bitsw = 0;
        for (int c = 0; c < 64; ++c) {
            const int q = w * 64 + c;
            if (q > r && q < n && box_iou_gt(a, b4[s_idx[q]], thr)) bitsw |= 1ull << c;
        }
        supp[t] = bitsw;
    }
    __syncthreads();
    // sequential scan by one warp: lane l owns words l, l + 32, ... of the removed mask
    if (tid < 32) {
        constexpr int MAXW = 8;  // n <= 16384
        unsigned long long removed[MAXW];
#pragma unroll
        for (int k = 0; k < MAXW; ++k) removed[k] = 0;
        int kept = 0;
        for (int r = 0; r < n; ++r) {
            const int w = r >> 6;
            unsigned long long mine = 0;
#pragma unroll
            for (int k = 0; k < MAXW; ++k)
                if (w == k * 32 + tid) mine = removed[k];
            const unsigned long long rw = __shfl_sync(0xffffffffu, mine, w & 31);
            const bool dead = (rw >> (r & 63)) & 1ull;
            if (tid == 0) keep[s_idx[r]] = dead ? 0 : 1;
            if (!dead) {
                ++kept;
#pragma unroll
                for (int k = 0; k < MAXW; ++k) {
                    const int ww = k * 32 + tid;
                    if (ww < words) removed[k] |= supp[(int64_t)r * words + ww];
                }
            }
        }
        if (tid == 0) *n_keep = kept;
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int64_t marsb200_rle_workspace_bytes(int64_t n, int H, int W) {
    if (n <= 0 || H <= 0 || W <= 0) return 0;
    return n * (int64_t)W * (H / 32) * 4;
}

int marsb200_rle_decode(const int32_t* counts, const int64_t* offsets, int64_t n, int H, int W, uint32_t* bits,
                        void* workspace, int64_t workspace_bytes, int32_t* status, void* stream) {
    MARS_REQUIRE(counts && offsets && bits && workspace && status, "null pointer");
    MARS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0, "shape");
    if ((H & 31) || (W & 31))
        return fail(MARSB200_ERR_UNSUPPORTED, "%s: H and W must be multiples of 32 (%lld x %lld)", "marsb200_rle_decode", H, W);
    MARS_REQUIRE(workspace_bytes >= marsb200_rle_workspace_bytes(n, H, W), "workspace too small");
    cudaStream_t s = as_stream(stream);
    const int64_t hw = (int64_t)H * W, col_words = (int64_t)W * (H / 32);
    const int64_t wpm = marsb200_words_per_mask(hw);
    MARS_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)(n * col_words * 4), s));
    MARS_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    if (wpm * 32 != hw) MARS_CUDA_OK(cudaMemsetAsync(bits, 0, (size_t)(n * wpm * 4), s));  // padding words
    rle_fill_kernel<<<(unsigned)n, RLE_THREADS, 0, s>>>(counts, offsets, hw, col_words, (uint32_t*)workspace, status);
    MARS_LAUNCH_OK();
    const unsigned blocks = (unsigned)(ceil_div(H / 32, BT_TILES) * ceil_div(W / 32, BT_TILES));
    bit_transpose_kernel<<<dim3(blocks, (unsigned)n), 256, 0, s>>>((const uint32_t*)workspace, H, W, wpm, bits);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_mask_boxes(const uint32_t* bits, int64_t n, int H, int W, int32_t* boxes, void* stream) {
    MARS_REQUIRE(bits && boxes, "null pointer");
    MARS_REQUIRE(n > 0 && H > 0 && W > 0, "shape");
    mask_boxes_kernel<<<(unsigned)n, 256, 0, as_stream(stream)>>>(bits, marsb200_words_per_mask((int64_t)H * W), H, W, boxes);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_stability_score(const float* logits, int64_t n, int64_t HW, float mask_threshold, float threshold_offset,
                             float* out, int32_t* counts, void* stream) {
    MARS_REQUIRE(logits && out && counts, "null pointer");
    MARS_REQUIRE(n > 0 && n <= 65535 && HW > 0, "shape");
    cudaStream_t s = as_stream(stream);
    MARS_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)(n * 2 * 4), s));
    const unsigned gx = (unsigned)std::min<int64_t>(ceil_div64(HW, 256 * 16), 1024);
    stability_counts_kernel<<<dim3(gx, (unsigned)n), 256, 0, s>>>(logits, HW, mask_threshold + threshold_offset,
                                                                  mask_threshold - threshold_offset, counts);
    MARS_LAUNCH_OK();
    stability_finish_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(counts, n, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int64_t marsb200_box_nms_workspace_bytes(int n) {
    if (n <= 0) return 0;
    return (int64_t)n * ((n + 63) / 64) * 8;
}

int marsb200_box_nms(const float* boxes, const float* scores, int n, float iou_threshold, int32_t* order, uint8_t* keep,
                     int32_t* n_keep, void* workspace, int64_t workspace_bytes, void* stream) {
    MARS_REQUIRE(boxes && scores && order && keep && n_keep && workspace, "null pointer");
    MARS_REQUIRE(n > 0 && n <= 16384, "1 <= n <= 16384");
    MARS_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "alignment");
    MARS_REQUIRE(workspace_bytes >= marsb200_box_nms_workspace_bytes(n), "workspace too small");
    int n_pow2 = 1;
    while (n_pow2 < n) n_pow2 <<= 1;
    const size_t smem = (size_t)n_pow2 * 8;
    MARS_CUDA_OK(cudaFuncSetAttribute(box_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    box_nms_kernel<<<1, NMS_THREADS, smem, as_stream(stream)>>>(boxes, scores, n, n_pow2, iou_threshold, order,
                                                                (unsigned long long*)workspace, keep, n_keep);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

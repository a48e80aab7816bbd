// Score fusion, stable ranking, greedy IoU-NMS and merge selection (one CTA per episode), plus
// the AlphaCLIP cosine scores.
#include <cuda_fp16.h>

#include "common.cuh"

namespace marsb200 {

#ifdef MARSB200_FUSE_PROFILE
#define FP_TIC() long long fp_last = clock64(); long long fp_acc[12] = {0}; int fp_k = 0
#define FP_LAP() do { const long long t = clock64(); fp_acc[fp_k++] = t - fp_last; fp_last = t; } while (0)
#define FP_PRINT() do { if (threadIdx.x == 0 && blockIdx.x == 0) printf("fuse_rank cycles: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", fp_acc[0], fp_acc[1], fp_acc[2], fp_acc[3], fp_acc[4], fp_acc[5], fp_acc[6], fp_acc[7], fp_acc[8], fp_acc[9]); } while (0)
#else
#define FP_TIC() do { } while (0)
#define FP_LAP() do { } while (0)
#define FP_PRINT() do { } while (0)
#endif

__global__ void clip_scores_kernel(const float* __restrict__ img, const float* __restrict__ txt, int64_t total, int P,
                                   int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wg >= total) return;
    const int64_t e = wg / P;
    const float* a = img + wg * D;
    const float* t = txt + e * D;
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) acc += (double)a[d] * (double)t[d];
    acc = warp_sum(acc);
    if (lane == 0) out[wg] = (float)acc;
}

// float16 features (the reference's AlphaCLIP runs in half precision on a GPU, FilteringMergingModule.py:189,195): the dot
// product is accumulated in fp32 and rounded to float16 once, like a half-precision matmul; `out` holds that float16 value.
__global__ void clip_scores_f16_kernel(const __half* __restrict__ img, const __half* __restrict__ txt, int64_t total, int P,
                                       int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wg >= total) return;
    const int64_t e = wg / P;
    const __half* a = img + wg * D;
    const __half* t = txt + e * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(__half2float(a[d]), __half2float(t[d]), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[wg] = __half2float(__float2half_rn(acc));
}

// one float16 operation the way NumPy evaluates it: operands are float16 values, the operation runs in float32 and the
// result is rounded to float16 (round to nearest even)
__device__ __forceinline__ float h_rn(float x) { return __half2float(__float2half_rn(x)); }

// one CTA of FUSE_THREADS threads ranks an episode: the kernel is bound by what a single SM can issue
constexpr int FUSE_THREADS = 1024;

// Pairwise suppression relation of one episode in PROPOSAL space: sup[i] bit j = IoU(i, j) > thr (j != i), from the
// intersection matrix (diagonal = area).  It does not depend on the ranking, so it is computed by a many-CTA kernel as
// soon as the intersections exist (marsb200_nms_bitmask) instead of by the single CTA that ranks the episode.  One warp
// per row i: every request is 32 consecutive entries of row i, balloted into one word; rows whose intersections with
// a whole 32-proposal block are all zero (the usual case) cost three instructions per word.
__global__ void __launch_bounds__(256) nms_bitmask_kernel(const int32_t* __restrict__ inter, int P, int nw, float nms_thr,
                                                          int64_t rows_total, uint32_t* __restrict__ sup) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // e * P + i
    if (row >= rows_total) return;
    const int64_t e = row / P;
    const int i = (int)(row - e * P);
    const int32_t* im = inter + e * (int64_t)P * P;
    const int32_t* ri = im + (int64_t)i * P;
    const int ai = ri[i];
    uint32_t* out = sup + row * nw;
    for (int w0 = 0; w0 < nw; w0 += 8) {
        int in[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = (w0 + q) * 32 + lane;
            in[q] = (w0 + q < nw && j < P && j != i) ? ri[j] : 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (w0 + q >= nw) break;  // warp-uniform
            uint32_t bitsw = 0;
            if (__any_sync(0xffffffffu, in[q] != 0)) {
                const int j = (w0 + q) * 32 + lane;
                bool over = false;
                if (in[q] != 0) {
                    const int un = ai + im[(int64_t)j * P + j] - in[q];
                    // the IEEE division only where the quotient can exceed the threshold at all: 2 * in > thr * un is
                    // implied by in / un > thr with a factor-two margin (float rounding is ~1e-7)
                    if (un > 0 && 2.f * (float)in[q] > nms_thr * (float)un) over = __fdiv_rn((float)in[q], (float)un) > nms_thr;
                }
                bitsw = __ballot_sync(0xffffffffu, over);
            }
            if (lane == 0) out[w0 + q] = bitsw;
        }
    }
}

// sort key order: higher score first, equal scores keep ascending proposal index (stable sort of the
// reference's sorted(reverse=True), FilteringMergingModule.py:138)
__device__ __forceinline__ bool ranks_before(double sa, int ia, double sb, int ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(FUSE_THREADS) fuse_rank_kernel(
    const double* __restrict__ emd, const float* __restrict__ clip, const int32_t* __restrict__ pooled_count,
    const float* __restrict__ sum_vva, const float* __restrict__ sum_vta, const int32_t* __restrict__ union_count,
    const int32_t* __restrict__ inter, const uint32_t* __restrict__ nms_bits, int P, int n2, int use_bitmask, int clip_f16, double alpha, double static_thr, double dynamic_thr,
    float nms_thr, double* __restrict__ scores, int32_t* __restrict__ order, uint8_t* __restrict__ flags,
    int32_t* __restrict__ summary, uint8_t* __restrict__ record, int64_t record_stride) {
    extern __shared__ unsigned char smem_raw[];
    double* s_key = reinterpret_cast<double*>(smem_raw);               // n2
    int* s_idx = reinterpret_cast<int*>(s_key + n2);                   // n2
    int* s_rank = s_idx + n2;                                          // P: proposal index -> rank
    int* s_area = s_rank + P;                                          // P
    unsigned char* s_removed = reinterpret_cast<unsigned char*>(s_area + P);  // P (+ pad to 4)
    uint32_t* s_sup = use_bitmask ? reinterpret_cast<uint32_t*>(s_removed + ((P + 3) & ~3)) : nullptr;  // (P + 8) * ceil(P/32)
    __shared__ double s_scratch_d[FUSE_THREADS / 32], s_scratch_d2[FUSE_THREADS / 32];
    __shared__ float s_scratch_f[FUSE_THREADS / 32], s_scratch_f2[FUSE_THREADS / 32];
    __shared__ int s_counts[2];
    __shared__ int s_nonfinite;  // proposals whose fused score is NaN / inf (e.g. a NaN EMD input): reported in summary[3]
    if (threadIdx.x == 0) s_nonfinite = 0;
    FP_TIC();

    const int64_t e = blockIdx.x;
    const int tid = threadIdx.x;
    // optional result record of the episode (the row of the all-gather table, episodes.py):
    // order int32[P] | score float32[P] | flags uint8[pad4(P)] | summary int32[4]
    int32_t* rec_order = record ? reinterpret_cast<int32_t*>(record + e * record_stride) : nullptr;
    float* rec_score = record ? reinterpret_cast<float*>(record + e * record_stride + 4 * (int64_t)P) : nullptr;
    uint8_t* rec_flags = record ? record + e * record_stride + 8 * (int64_t)P : nullptr;
    int32_t* rec_summary = record ? reinterpret_cast<int32_t*>(record + e * record_stride + 8 * (int64_t)P + ((P + 3) & ~3)) : nullptr;
    emd += e * P;
    clip += e * P;
    pooled_count += e * P;
    sum_vva += e * P;
    sum_vta += e * P;

    // ---- min / max of the two globally normalised scores
    double emin = INFINITY, emax = -INFINITY;
    float cmin = INFINITY, cmax = -INFINITY;
    for (int p = tid; p < P; p += FUSE_THREADS) {
        const double v = emd[p];
        const float c = clip[p];
        emin = fmin(emin, v);
        emax = fmax(emax, v);
        cmin = fminf(cmin, c);
        cmax = fmaxf(cmax, c);
    }
    {
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
            emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
        if (lane == 0) {
            s_scratch_d[warp] = emin;
            s_scratch_d2[warp] = emax;
            s_scratch_f[warp] = cmin;
            s_scratch_f2[warp] = cmax;
        }
        __syncthreads();
        emin = s_scratch_d[lane];
        emax = s_scratch_d2[lane];
        cmin = s_scratch_f[lane];
        cmax = s_scratch_f2[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
            emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
    }
    FP_LAP();
    const double e_den = 1e-7 + emax - emin;         // float64, FilteringMergingModule.py:131
    // evaluated in the feature dtype, :132 - float32, or float16 (`1e-7` is a weak Python scalar: 1.19e-7 as a half)
    const float c_den = clip_f16 ? h_rn(h_rn(h_rn(1e-7f) + cmax) - cmin) : (1e-7f + cmax) - cmin;
    const double u_den = 1e-7 + (double)union_count[e];

    // ---- fused score per proposal
    for (int p = tid; p < n2; p += FUSE_THREADS) {
        if (p < P) {
            const double cnt = (double)pooled_count[p];
            const double cov = cnt / u_den;
            const double avv = (double)sum_vva[p] / (1e-7 + cnt);
            const double avt = (double)sum_vta[p] / (1e-7 + cnt);
            const double pvv = alpha * avv + (1.0 - alpha) * cov;
            const double pvt = alpha * avt + (1.0 - alpha) * cov;
            const double en = (emd[p] - emin) / e_den;
            double sc;
            if (clip_f16) {
                // NumPy's types at :136 - Python float + float16 array: the float is cast to float16 and the sum is a
                // float16 operation; adding the np.float64 pvv / pvt then promotes to float64
                const float cn = h_rn(__fdiv_rn(h_rn(clip[p] - cmin), c_den));
                const float first = h_rn(__half2float(__double2half(en)) + cn);
                sc = (((double)first + pvv) + pvt) / 4.0;
            } else {
                const float cn = __fdiv_rn(clip[p] - cmin, c_den);
                sc = (((en + (double)cn) + pvv) + pvt) / 4.0;
            }
            scores[e * P + p] = sc;
            if (rec_score) rec_score[p] = (float)sc;
            const bool finite = isfinite(sc);
            if (!finite) atomicAdd(&s_nonfinite, 1);
            s_key[p] = finite ? sc : -INFINITY;  // a NaN key would make the sort inconsistent: such proposals rank last
            s_idx[p] = p;
        } else {
            s_key[p] = -INFINITY;
            s_idx[p] = 0x7fffffff;
        }
    }
    __syncthreads();

    FP_LAP();
    // ---- rank order.  P <= 1024: every proposal counts the proposals that rank before it (no barriers: `parts` adjacent
    // lanes split the scan of the keys); larger P: bitonic sort of (key, idx)
    if (n2 <= FUSE_THREADS) {
        int parts = FUSE_THREADS / n2;  // a power of two
        if (parts > 32) parts = 32;
        const int p = tid / parts, part = tid - p * parts;
        const bool live = p < P;
        const double mine = live ? s_key[p] : 0.0;
        int before = 0;
        if (live) {
            // q ranks before p iff key_q > key_p, or the keys tie and q < p.  The `parts` lanes of a proposal take
            // interleaved q (adjacent shared-memory words: one conflict-free request per step for the whole warp)
#pragma unroll 4
            for (int q = part; q < P; q += parts) {
                const double kq = s_key[q];
                const bool gt = kq > mine, ge = kq >= mine;
                before += (q < p ? ge : gt) ? 1 : 0;
            }
        }
        FP_LAP();
        for (int o = parts >> 1; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        __syncthreads();  // every key has been read
        FP_LAP();
        if (live && part == 0) {
            s_key[before] = mine;
            s_idx[before] = p;
        }
        __syncthreads();
    } else {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += FUSE_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;  // this sub-sequence ends in rank order
                    const double ki = s_key[i], kl = s_key[l];
                    const int ii = s_idx[i], il = s_idx[l];
                    const bool l_first = ranks_before(kl, il, ki, ii);
                    if (l_first == up) {
                        s_key[i] = kl;
                        s_key[l] = ki;
                        s_idx[i] = il;
                        s_idx[l] = ii;
                    }
                }
            }
            __syncthreads();
        }
    }
    }
    for (int r = tid; r < P; r += FUSE_THREADS) {
        const int p = s_idx[r];
        order[e * P + r] = p;
        if (rec_order) rec_order[r] = p;
        s_rank[p] = r;
        s_removed[r] = 0;  // indexed by proposal below; cleared for all P entries here
    }
    __syncthreads();

    FP_LAP();
    // ---- greedy IoU-NMS in rank order (builder-defined; torchvision nms semantics)
    const bool do_nms = inter != nullptr && nms_thr >= 0.f;
    if (do_nms) {
        const int32_t* im = inter + e * (int64_t)P * P;
        for (int p = tid; p < P; p += FUSE_THREADS) s_area[p] = im[(int64_t)p * P + p];
        __syncthreads();
        FP_LAP();
        if (s_sup != nullptr) {
            const int nw = (P + 31) >> 5;
            const int lane = tid & 31;
            // the scan reads up to seven ranks past the end: zero rows suppress nothing whatever index is read with them
            for (int k = tid; k < 8 * nw; k += FUSE_THREADS) s_sup[P * nw + k] = 0u;
            if (nms_bits != nullptr) {
                // (a) the relation was computed ahead of the ranking (marsb200_nms_bitmask): stage its rows in RANK order -
                // row r of the shared copy is the row of proposal order[r] - so the scan below reads them sequentially
                const uint32_t* gb = nms_bits + e * (int64_t)P * nw;
                for (int idx = tid; idx < P * nw; idx += FUSE_THREADS) {
                    const int r = idx / nw, w = idx - r * nw;
                    s_sup[idx] = gb[(int64_t)s_idx[r] * nw + w];
                }
            } else {
                // (a') built here: one warp per (rank, word), 32 consecutive entries of the proposal's row per request
                constexpr int WPI = 8;  // words per warp iteration: that many independent row reads in flight per lane
                for (int idx0 = (tid >> 5) * WPI; idx0 < P * nw; idx0 += (FUSE_THREADS / 32) * WPI) {
                    int in[WPI], ii[WPI], jj[WPI];
#pragma unroll
                    for (int q = 0; q < WPI; ++q) {
                        const int idx = idx0 + q;
                        const int r = idx / nw;
                        ii[q] = idx < P * nw ? s_idx[r] : 0;
                        jj[q] = (idx - r * nw) * 32 + lane;
                        in[q] = (idx < P * nw && jj[q] < P) ? im[(int64_t)ii[q] * P + jj[q]] : 0;
                    }
#pragma unroll
                    for (int q = 0; q < WPI; ++q) {
                        const int idx = idx0 + q;
                        if (idx >= P * nw) break;  // warp-uniform
                        bool over = false;
                        if (jj[q] < P && jj[q] != ii[q] && in[q] != 0) {
                            const int un = s_area[ii[q]] + s_area[jj[q]] - in[q];
                            if (un > 0 && 2.f * (float)in[q] > nms_thr * (float)un)
                                over = __fdiv_rn((float)in[q], (float)un) > nms_thr;
                        }
                        const uint32_t bitsw = __ballot_sync(0xffffffffu, over);
                        if (lane == 0) s_sup[idx] = bitsw;
                    }
                }
            }
            __syncthreads();
            FP_LAP();
            // (b) one warp walks the ranks; lane w owns word w of the removed mask in proposal space (P <= 1024).  The
            // relation is symmetric, so no "later ranks only" filter is needed: a kept proposal never overlaps an earlier
            // kept one, and the bits of earlier removed ones are set already.
            if (tid < 32) {
                // One warp runs alone here, so a rank costs what its dependent chain costs: shuffle -> shift -> mask ->
                // or.  Rows and indices are padded with four zero entries (no bounds tests) and fetched four ranks ahead.
                uint32_t removed = 0;
                const int col = tid < nw ? tid : 0;
                const uint32_t live = tid < nw ? 0xffffffffu : 0u;
                const uint32_t* rowp = s_sup + col;
                const int* idxp = s_idx;
                int i0 = idxp[0], i1 = idxp[1], i2 = idxp[2], i3 = idxp[3];
                uint32_t u0 = rowp[0], u1 = rowp[nw], u2 = rowp[2 * nw], u3 = rowp[3 * nw];
                for (int r0 = 0; r0 < P; r0 += 4) {
                    idxp += 4;
                    rowp += 4 * nw;
                    const int j0 = idxp[0], j1 = idxp[1], j2 = idxp[2], j3 = idxp[3];
                    const uint32_t v0 = rowp[0], v1 = rowp[nw], v2 = rowp[2 * nw], v3 = rowp[3 * nw];
                    uint32_t word;
                    word = __shfl_sync(0xffffffffu, removed, i0 >> 5);
                    removed |= u0 & (((word >> (i0 & 31)) & 1u) - 1u);
                    word = __shfl_sync(0xffffffffu, removed, i1 >> 5);
                    removed |= u1 & (((word >> (i1 & 31)) & 1u) - 1u);
                    word = __shfl_sync(0xffffffffu, removed, i2 >> 5);
                    removed |= u2 & (((word >> (i2 & 31)) & 1u) - 1u);
                    word = __shfl_sync(0xffffffffu, removed, i3 >> 5);
                    removed |= u3 & (((word >> (i3 & 31)) & 1u) - 1u);
                    i0 = j0; i1 = j1; i2 = j2; i3 = j3;
                    u0 = v0; u1 = v1; u2 = v2; u3 = v3;
                }
                removed &= live;
#pragma unroll 1
                for (int b = 0; b < 32; ++b) {
                    const int p = tid * 32 + b;
                    if (tid < nw && p < P) s_removed[p] = (removed >> b) & 1u;
                }
            }
            __syncthreads();
        } else {
            for (int r = 0; r < P; ++r) {
                const int i = s_idx[r];
                if (s_removed[i]) continue;  // uniform: shared state is stable between barriers
                const int ai = s_area[i];
                const int32_t* row = im + (int64_t)i * P;
                for (int j = tid; j < P; j += FUSE_THREADS) {
                    if (s_rank[j] > r && !s_removed[j]) {
                        const int in = row[j];
                        const int un = ai + s_area[j] - in;
                        const float iou = un > 0 ? __fdiv_rn((float)in, (float)un) : 0.f;
                        if (iou > nms_thr) s_removed[j] = 1;
                    }
                }
                __syncthreads();
            }
        }
    }

    FP_LAP();
    // ---- merge selection (FilteringMergingModule.py:213-217) and outputs
    if (tid == 0) s_counts[0] = s_counts[1] = 0;
    __syncthreads();
    const double top = s_key[0];
    const double bound = (top < static_thr) ? dynamic_thr * top : static_thr;
    int kept = 0, selected = 0;
    for (int p = tid; p < P; p += FUSE_THREADS) {
        const bool keep = !do_nms || !s_removed[p];
        const bool sel = keep && (s_key[s_rank[p]] >= bound);
        flags[e * P + p] = (keep ? 1 : 0) | (sel ? 2 : 0);
        if (rec_flags) rec_flags[p] = (keep ? 1 : 0) | (sel ? 2 : 0);
        kept += keep;
        selected += sel;
    }
    kept = warp_sum(kept);
    selected = warp_sum(selected);
    if ((tid & 31) == 0) {
        atomicAdd(&s_counts[0], kept);
        atomicAdd(&s_counts[1], selected);
    }
    __syncthreads();
    if (tid == 0) {
        summary[e * 4 + 0] = s_counts[0];
        summary[e * 4 + 1] = s_counts[1];
        summary[e * 4 + 2] = s_idx[0];
        summary[e * 4 + 3] = s_nonfinite;
        if (rec_summary) {
            rec_summary[0] = s_counts[0];
            rec_summary[1] = s_counts[1];
            rec_summary[2] = s_idx[0];
            rec_summary[3] = s_nonfinite;
        }
    }
    FP_LAP();
    FP_PRINT();
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int marsb200_clip_scores(const float* img, const float* txt, int E, int P, int D, float* out, void* stream) {
    MARS_REQUIRE(img && txt && out, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && D > 0, "shape");
    const int64_t total = (int64_t)E * P;
    clip_scores_kernel<<<(unsigned)ceil_div64(total, 8), 256, 0, as_stream(stream)>>>(img, txt, total, P, D, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_clip_scores_f16(const void* img, const void* txt, int E, int P, int D, float* out, void* stream) {
    MARS_REQUIRE(img && txt && out, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && D > 0, "shape");
    const int64_t total = (int64_t)E * P;
    clip_scores_f16_kernel<<<(unsigned)ceil_div64(total, 8), 256, 0, as_stream(stream)>>>(
        static_cast<const __half*>(img), static_cast<const __half*>(txt), total, P, D, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int64_t marsb200_record_bytes(int P) { return P <= 0 ? 0 : 8 * (int64_t)P + ((P + 3) & ~3) + 16; }

int marsb200_nms_bitmask(const int32_t* inter, int E, int P, float nms_iou_threshold, uint32_t* nms_bits, void* stream) {
    MARS_REQUIRE(inter && nms_bits, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && P <= 8192, "shape (P <= 8192)");
    MARS_REQUIRE(nms_iou_threshold >= 0.f, "nms_iou_threshold >= 0");
    const int nw = (P + 31) / 32;
    const int64_t rows = (int64_t)E * P;
    nms_bitmask_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, as_stream(stream)>>>(inter, P, nw, nms_iou_threshold, rows, nms_bits);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_fuse_rank(const double* emd, const float* clip, const int32_t* pooled_count, const float* sum_vva,
                       const float* sum_vta, const int32_t* union_count, const int32_t* inter, const uint32_t* nms_bits,
                       int E, int P, double alpha, double static_threshold, double dynamic_threshold,
                       float nms_iou_threshold, int clip_f16, double* scores, int32_t* order, uint8_t* flags,
                       int32_t* summary, uint8_t* record, int64_t record_stride, void* stream) {
    MARS_REQUIRE(emd && clip && pooled_count && sum_vva && sum_vta && union_count, "null input");
    MARS_REQUIRE(scores && order && flags && summary, "null output");
    MARS_REQUIRE(E > 0 && P > 0 && P <= 8192, "shape (P <= 8192)");
    MARS_REQUIRE(!record || (record_stride >= marsb200_record_bytes(P) && record_stride % 4 == 0 &&
                             (reinterpret_cast<uintptr_t>(record) & 3) == 0), "record row: >= marsb200_record_bytes(P), 4-byte aligned");
    MARS_REQUIRE(!nms_bits || inter, "nms_bits needs inter (the areas are its diagonal)");
    int n2 = 1;
    while (n2 < P) n2 <<= 1;
    size_t smem = (size_t)n2 * (sizeof(double) + sizeof(int)) + (size_t)P * (2 * sizeof(int) + 1) + 16;
    // suppression bitmask (P * ceil(P/32) words) when it fits in shared memory
    const bool nms = inter != nullptr && nms_iou_threshold >= 0.f;
    const size_t sup_bytes = (size_t)(P + 8) * ((P + 31) / 32) * 4;  // + eight zero rows behind the last rank
    const int use_bitmask = nms && P <= 1024 && smem + sup_bytes <= 200 * 1024;
    if (use_bitmask) smem += sup_bytes;
    if (smem > 48 * 1024)
        MARS_CUDA_OK(cudaFuncSetAttribute(fuse_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fuse_rank_kernel<<<E, FUSE_THREADS, smem, as_stream(stream)>>>(emd, clip, pooled_count, sum_vva, sum_vta,
                                                                   union_count, inter, use_bitmask ? nms_bits : nullptr, P, n2, use_bitmask, clip_f16 ? 1 : 0, alpha, static_threshold,
                                                                   dynamic_threshold, nms_iou_threshold, scores, order,
                                                                   flags, summary, record, record_stride);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

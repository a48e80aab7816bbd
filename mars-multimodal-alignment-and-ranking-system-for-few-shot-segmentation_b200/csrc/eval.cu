// Evaluation accumulators: the per-class intersection / union buffers of AverageMeter
// (mars/utils/logger.py:42-44, 61-78) kept on the device as exact 64-bit integer pixel counts, so a whole
// evaluation run needs no host synchronisation and ranks can be combined with one all-reduce.
#include "common.cuh"

namespace marsb200 {

// areas [n, 4] = {inter_bg, inter_fg, union_bg, union_fg} (marsb200_eval_areas); class_id [n];
// inter_buf / union_buf [2, nclass] += ... (AverageMeter.update: index_add_ along dim 1)
__global__ void __launch_bounds__(256) eval_accumulate_kernel(const int32_t* __restrict__ areas,
                                                               const int64_t* __restrict__ class_id, int64_t n, int nclass,
                                                               unsigned long long* __restrict__ inter_buf,
                                                               unsigned long long* __restrict__ union_buf,
                                                               int* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = class_id[i];
    if (c < 0 || c >= nclass) {
        atomicExch(status, 1);
        return;
    }
    atomicAdd(inter_buf + c, (unsigned long long)areas[4 * i + 0]);
    atomicAdd(inter_buf + nclass + c, (unsigned long long)areas[4 * i + 1]);
    atomicAdd(union_buf + c, (unsigned long long)areas[4 * i + 2]);
    atomicAdd(union_buf + nclass + c, (unsigned long long)areas[4 * i + 3]);
}

// AverageMeter.compute_iou (logger.py:69-78): iou = inter / max(union, 1) on the classes of interest;
// out[0] = mIoU = mean(iou_fg) * 100, out[1] = FB-IoU = mean over {bg, fg} of (sum inter / sum union) * 100,
// out[2 + k] = iou_fg of the k-th class of interest.  One block.
__global__ void __launch_bounds__(256) eval_iou_kernel(const unsigned long long* __restrict__ inter_buf,
                                                        const unsigned long long* __restrict__ union_buf, int nclass,
                                                        const int64_t* __restrict__ interest, int k, double* __restrict__ out) {
    __shared__ double s_iou[8];
    __shared__ unsigned long long s_sum[4][8];
    double iou_sum = 0.0;
    unsigned long long si[2] = {0, 0}, su[2] = {0, 0};
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const int64_t c = interest[t];
        const unsigned long long ifg = inter_buf[nclass + c], ufg = union_buf[nclass + c];
        const double iou = (double)ifg / (double)(ufg > 1 ? ufg : 1);
        out[2 + t] = iou;
        iou_sum += iou;
        si[0] += inter_buf[c];
        si[1] += ifg;
        su[0] += union_buf[c];
        su[1] += ufg;
    }
    iou_sum = warp_sum(iou_sum);
    for (int o = 16; o > 0; o >>= 1) {
        si[0] += __shfl_xor_sync(0xffffffffu, si[0], o);
        si[1] += __shfl_xor_sync(0xffffffffu, si[1], o);
        su[0] += __shfl_xor_sync(0xffffffffu, su[0], o);
        su[1] += __shfl_xor_sync(0xffffffffu, su[1], o);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_iou[warp] = iou_sum;
        s_sum[0][warp] = si[0];
        s_sum[1][warp] = si[1];
        s_sum[2][warp] = su[0];
        s_sum[3][warp] = su[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        unsigned long long a[4] = {0, 0, 0, 0};
        for (int w = 0; w < 8; ++w) {
            tot += s_iou[w];
            for (int q = 0; q < 4; ++q) a[q] += s_sum[q][w];
        }
        out[0] = k > 0 ? tot / k * 100.0 : nan("");
        out[1] = ((double)a[0] / (double)a[2] + (double)a[1] / (double)a[3]) / 2.0 * 100.0;  // 0/0 -> NaN like the reference
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" {

int marsb200_eval_accumulate(const int32_t* areas, const int64_t* class_id, int64_t n, int nclass, int64_t* inter_buf,
                             int64_t* union_buf, int32_t* status, void* stream) {
    MARS_REQUIRE(areas && class_id && inter_buf && union_buf && status, "null pointer");
    MARS_REQUIRE(n > 0 && nclass > 0, "shape");
    eval_accumulate_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(
        areas, class_id, n, nclass, reinterpret_cast<unsigned long long*>(inter_buf),
        reinterpret_cast<unsigned long long*>(union_buf), status);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

int marsb200_eval_iou(const int64_t* inter_buf, const int64_t* union_buf, int nclass, const int64_t* interest, int k,
                      double* out, void* stream) {
    MARS_REQUIRE(inter_buf && union_buf && interest && out, "null pointer");
    MARS_REQUIRE(nclass > 0 && k > 0, "shape");
    eval_iou_kernel<<<1, 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(inter_buf),
                                                      reinterpret_cast<const unsigned long long*>(union_buf), nclass,
                                                      interest, k, out);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"

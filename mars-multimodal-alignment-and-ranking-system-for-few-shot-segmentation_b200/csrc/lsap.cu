// Exact rectangular linear-sum assignment on the device: the solver behind Matcher's
// bidirectional patch matching, scipy.optimize.linear_sum_assignment(S, maximize=True) at
// matcher/Matcher.py:449-450 (forward: fg support rows x query patches) and :471-472 (reverse).
// Shortest augmenting paths (Jonker-Volgenant) with dense reduced costs in float64: one CTA per problem,
// block-wide arg-min over the sinks, parallel relaxation, sequential augmentation.  The smaller side is
// always the set of sources, like scipy's rectangular solver every source gets assigned.
#include "common.cuh"

namespace marsb200 {

constexpr int LSAP_THREADS = 512;
constexpr double LSAP_INF = 1e300;

// per source: u, dsrc (f64), id / assigned sink / reached list (i32); per sink: v, dist (f64), id / predecessor /
// assigned source (u16) and a scanned flag: 23 bytes, so that the 5-shot reverse matching (<= 1369 x 6845) fits
__host__ __device__ inline size_t lsap_smem_bytes(int t_cap, int m_cap) {
    return (size_t)t_cap * (2 * 8 + 3 * 4) + (size_t)m_cap * (2 * 8 + 3 * 2 + 1) + 64;
}
constexpr unsigned short LSAP_NONE = 0xffff;

__device__ inline void lsap_argmin(const double* dist, const unsigned char* scanned, int M, double& best, int& best_j,
                                   double* s_val, int* s_idx) {
    double v = LSAP_INF;
    int j = 0x7fffffff;
    for (int t = threadIdx.x; t < M; t += LSAP_THREADS) {
        const double k = scanned[t] ? LSAP_INF : dist[t];
        if (k < v) {
            v = k;
            j = t;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        if (ov < v || (ov == v && oj < j)) {
            v = ov;
            j = oj;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_val[threadIdx.x >> 5] = v;
        s_idx[threadIdx.x >> 5] = j;
    }
    __syncthreads();
    best = s_val[0];
    best_j = s_idx[0];
#pragma unroll
    for (int w = 1; w < LSAP_THREADS / 32; ++w) {
        if (s_val[w] < best || (s_val[w] == best && s_idx[w] < best_j)) {
            best = s_val[w];
            best_j = s_idx[w];
        }
    }
}

// sim [E, R, C]; row_sel [E, R] / col_sel [E, C] choose the participating rows / columns (null = all).
// row_to_col [E, R]: assigned column of each selected row or -1; objective [E]: sum of the assigned similarities.
__global__ void __launch_bounds__(LSAP_THREADS) lsap_kernel(const float* __restrict__ sim, const uint8_t* __restrict__ row_sel,
                                                             const uint8_t* __restrict__ col_sel, int R, int Ccols,
                                                             int maximize, int t_cap, int m_cap,
                                                             int32_t* __restrict__ row_to_col,
                                                             double* __restrict__ objective, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char lsap_smem_raw[];
    double* u = reinterpret_cast<double*>(lsap_smem_raw);  // [t_cap]
    double* dsrc = u + t_cap;                              // [t_cap]
    double* v = dsrc + t_cap;                              // [m_cap]
    double* dist = v + m_cap;                              // [m_cap]
    int* src_id = reinterpret_cast<int*>(dist + m_cap);    // [t_cap] row (or column) index of source i
    int* src_sink = src_id + t_cap;                        // [t_cap] sink assigned to source i
    int* list = src_sink + t_cap;                          // [t_cap] reached sources
    unsigned short* sink_id = reinterpret_cast<unsigned short*>(list + t_cap);  // [m_cap]
    unsigned short* pred_src = sink_id + m_cap;            // [m_cap]
    unsigned short* sink_src = pred_src + m_cap;           // [m_cap] source assigned to sink j, or LSAP_NONE
    unsigned char* scanned = reinterpret_cast<unsigned char*>(sink_src + m_cap);  // [m_cap]
    __shared__ double s_val[2][LSAP_THREADS / 32];
    __shared__ int s_idx[2][LSAP_THREADS / 32];
    __shared__ int s_nr, s_nc, s_nreached;
    const int tid = threadIdx.x;
    const int64_t e = blockIdx.x;
    const float* S = sim + e * (int64_t)R * Ccols;
    int32_t* out = row_to_col + e * R;

    // selected rows / columns (ascending); the lists land in the sink arrays first and are swapped below if needed
    for (int r = tid; r < R; r += LSAP_THREADS) out[r] = -1;
    if (tid == 0) {
        int nr = 0, nc = 0;
        for (int r = 0; r < R; ++r) nr += (!row_sel || row_sel[e * R + r]) ? 1 : 0;
        for (int c = 0; c < Ccols; ++c) nc += (!col_sel || col_sel[e * Ccols + c]) ? 1 : 0;
        s_nr = nr;
        s_nc = nc;
    }
    __syncthreads();
    const int nr = s_nr, nc = s_nc;
    const bool rows_are_sources = nr <= nc;
    const int T = rows_are_sources ? nr : nc, M = rows_are_sources ? nc : nr;
    if (T == 0) {
        if (tid == 0) objective[e] = 0.0;
        return;
    }
    if (T > t_cap || M > m_cap) {
        if (tid == 0) {
            objective[e] = nan("");
            atomicMax(status, max(T, M));
        }
        return;
    }
    if (tid == 0) {
        int a = 0, b = 0;
        for (int r = 0; r < R; ++r)
            if (!row_sel || row_sel[e * R + r]) {
                if (rows_are_sources) src_id[a++] = r;
                else sink_id[a++] = (unsigned short)r;
            }
        for (int c = 0; c < Ccols; ++c)
            if (!col_sel || col_sel[e * Ccols + c]) {
                if (rows_are_sources) sink_id[b++] = (unsigned short)c;
                else src_id[b++] = c;
            }
    }
    __syncthreads();
    const double sign = maximize ? -1.0 : 1.0;
    auto cost = [&](int i, int j) -> double {
        const int r = rows_are_sources ? src_id[i] : sink_id[j];
        const int c = rows_are_sources ? sink_id[j] : src_id[i];
        return sign * (double)S[(int64_t)r * Ccols + c];
    };

    // duals: sinks may stay unassigned (inequality), so their dual starts at 0 and only decreases; the sources are
    // all assigned (equality), so u_i = min_j c_ij makes every reduced cost non-negative
    for (int i = tid >> 5; i < T; i += LSAP_THREADS / 32) {
        double mn = LSAP_INF;
        for (int j = tid & 31; j < M; j += 32) mn = fmin(mn, cost(i, j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if ((tid & 31) == 0) {
            u[i] = mn;
            src_sink[i] = -1;
        }
    }
    for (int j = tid; j < M; j += LSAP_THREADS) {
        v[j] = 0.0;
        sink_src[j] = LSAP_NONE;
    }
    __syncthreads();

    for (int r = 0; r < T; ++r) {
        const double ur = u[r];
        for (int j = tid; j < M; j += LSAP_THREADS) {
            const double d = cost(r, j) - ur - v[j];
            dist[j] = d;
            scanned[j] = 0;
            pred_src[j] = (unsigned short)r;
        }
        if (tid == 0) {
            dsrc[r] = 0.0;
            list[0] = r;
            s_nreached = 1;
        }
        __syncthreads();
        double D;
        int jstar;
        for (int step = 0;; ++step) {
            lsap_argmin(dist, scanned, M, D, jstar, s_val[step & 1], s_idx[step & 1]);
            const int i = sink_src[jstar] == LSAP_NONE ? -1 : (int)sink_src[jstar];  // only changes in the augmentation
            if (tid == 0) scanned[jstar] = 1;
            if (i < 0) break;
            if (tid == 0) {
                dsrc[i] = D;
                list[s_nreached++] = i;
            }
            const double base = D - u[i];
            __syncthreads();  // scanned[jstar] visible before the relaxation reads it
            for (int j = tid; j < M; j += LSAP_THREADS) {
                if (!scanned[j]) {
                    const double nd = base + cost(i, j) - v[j];
                    if (nd < dist[j]) {
                        dist[j] = nd;
                        pred_src[j] = (unsigned short)i;
                    }
                }
            }
            __syncthreads();
        }
        __syncthreads();
        const int nreached = s_nreached;
        for (int k = tid; k < nreached; k += LSAP_THREADS) {
            const int i = list[k];
            u[i] += D - dsrc[i];
        }
        for (int j = tid; j < M; j += LSAP_THREADS)
            if (scanned[j]) v[j] -= D - dist[j];
        __syncthreads();
        if (tid == 0) {  // flip the assignments along the path back to r
            int j = jstar;
            while (true) {
                const int i = pred_src[j];
                const int prev = src_sink[i];
                sink_src[j] = (unsigned short)i;
                src_sink[i] = j;
                if (i == r) break;
                j = prev;
            }
        }
        __syncthreads();
    }

    double acc = 0.0;
    for (int i = tid; i < T; i += LSAP_THREADS) {
        const int j = src_sink[i];
        const int r = rows_are_sources ? src_id[i] : sink_id[j];
        const int c = rows_are_sources ? sink_id[j] : src_id[i];
        out[r] = c;
        acc += (double)S[(int64_t)r * Ccols + c];
    }
    acc = warp_sum(acc);
    __syncthreads();
    if ((tid & 31) == 0) s_val[0][tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < LSAP_THREADS / 32; ++w) total += s_val[0][w];
        objective[e] = total;
    }
}

}  // namespace marsb200

using namespace marsb200;

extern "C" int marsb200_lsap(const float* sim, const uint8_t* row_sel, const uint8_t* col_sel, int E, int R, int C,
                             int maximize, int t_cap, int m_cap, int32_t* row_to_col, double* objective, int32_t* status,
                             void* stream) {
    MARS_REQUIRE(sim && row_to_col && objective && status, "null pointer");
    MARS_REQUIRE(E > 0 && R > 0 && C > 0 && R <= 65534 && C <= 65534, "shape (R, C <= 65534)");
    if (t_cap <= 0 || m_cap <= 0) {  // no bound on the selected counts given: size for the whole matrix
        t_cap = R < C ? R : C;
        m_cap = R < C ? C : R;
    }
    const size_t smem = lsap_smem_bytes(t_cap, m_cap);
    MARS_REQUIRE(smem <= 220 * 1024, "problem too large for the shared-memory state (28*min + 23*max bytes <= 220 KB)");
    cudaStream_t s = as_stream(stream);
    MARS_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MARS_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    lsap_kernel<<<E, LSAP_THREADS, smem, s>>>(sim, row_sel, col_sel, R, C, maximize, t_cap, m_cap, row_to_col, objective,
                                               status);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

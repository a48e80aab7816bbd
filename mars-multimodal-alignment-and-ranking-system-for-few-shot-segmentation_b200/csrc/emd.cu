// Exact earth mover's distance between the pooled support mask and every pooled proposal
// (SURVEY.md row A7 / 8f-1): the transportation LP the reference solves with POT's ot.emd2
// (FilteringMergingModule.py:142-169, matcher/Matcher.py:1187-1194), uniform marginals 1/T and 1/M,
// cost = C[fg support rows][proposal patches] in float64.
//
// One CTA per (episode, proposal).  Flows are kept in integer units of 1/(T*M) (supplies M, demands T),
// so the primal solution is exact; the solver is the successive-shortest-path (Hungarian-type)
// algorithm for the transportation problem on dense reduced costs: for each source with supply left,
// a Dijkstra over the sinks (block-wide argmin, parallel relaxation), dual update, augmentation along
// the alternating path.  The cost matrix is never gathered: c(i, j) = C[rows[i]][cols[j]] is read
// from the episode's [M_rows, N] cost matrix (L2 resident).  Optimal value is unique, so the result
// matches any exact LP solver up to float64 rounding.
#include "common.cuh"

namespace marsb200 {

#ifdef MARSB200_EMD_PROFILE
__device__ long long g_emd_prof[16];
#define EMD_TIC(t) long long t = clock64()
#define EMD_TOC(slot, t) do { if (threadIdx.x == 0) atomicAdd((unsigned long long*)&g_emd_prof[slot], (unsigned long long)(clock64() - (t))); } while (0)
#define EMD_COUNT(slot) do { if (threadIdx.x == 0) atomicAdd((unsigned long long*)&g_emd_prof[slot], 1ull); } while (0)
#else
#define EMD_TIC(t)
#define EMD_TOC(slot, t)
#define EMD_COUNT(slot)
#endif

constexpr int EMD_THREADS = 512;  // the relaxation gathers cost entries from L2: latency hidden by many threads
constexpr double EMD_INF = 1e300;
constexpr int EMD_INLINE = 6;        // a sink of a (near-)basic solution is fed by ~ (T + M) / M sources
constexpr int EMD_OVERFLOW = 255;

struct EmdSmem {
    double* u;       // [t_cap]  source duals
    double* dsrc;    // [t_cap]  distance at which a source was reached
    double* v;       // [n_cap]  sink duals
    double* dist;    // [n_cap]  tentative / final sink distances
    double* key;     // [n_cap]  dist for unscanned sinks, INF once scanned
    int* rows;       // [t_cap]  support row of source i
    int* pred_sink;  // [t_cap]  sink through which source i was reached (backward arc)
    int* supply;     // [t_cap]
    int* list;       // [t_cap]  reached sources (all) ...
    int* newlist;    // [t_cap]  ... and the ones reached in the current step
    int* cols;       // [n_cap]  patch index of sink j
    int* pred_src;   // [n_cap]  source that gave sink j its distance
    int* demand;     // [n_cap]
    unsigned char* reached;  // [t_cap]
    short* feeders;          // [n_cap][EMD_INLINE] sources with positive flow into sink j (when they fit)
    unsigned char* nfeed;    // [n_cap] how many, or EMD_OVERFLOW: scan the flow row in global memory instead
};

__host__ __device__ inline size_t emd_smem_bytes(int t_cap, int n_cap) {
    return (size_t)t_cap * (2 * 8 + 5 * 4 + 1) + (size_t)n_cap * (3 * 8 + 3 * 4 + 2 * EMD_INLINE + 1) + 64;
}

__device__ inline EmdSmem emd_carve(unsigned char* base, int t_cap, int n_cap) {
    EmdSmem s;
    double* d = reinterpret_cast<double*>(base);
    s.u = d;
    s.dsrc = s.u + t_cap;
    s.v = s.dsrc + t_cap;
    s.dist = s.v + n_cap;
    s.key = s.dist + n_cap;
    int* i = reinterpret_cast<int*>(s.key + n_cap);
    s.rows = i;
    s.pred_sink = s.rows + t_cap;
    s.supply = s.pred_sink + t_cap;
    s.list = s.supply + t_cap;
    s.newlist = s.list + t_cap;
    s.cols = s.newlist + t_cap;
    s.pred_src = s.cols + n_cap;
    s.demand = s.pred_src + n_cap;
    s.feeders = reinterpret_cast<short*>(s.demand + n_cap);
    s.reached = reinterpret_cast<unsigned char*>(s.feeders + (size_t)n_cap * EMD_INLINE);
    s.nfeed = s.reached + t_cap;
    return s;
}

// block-wide argmin of key[0..M) (ties -> smaller index); every thread gets the result
__device__ inline void block_argmin(const double* key, int M, double& best, int& best_j, double* s_val, int* s_idx) {
    // s_val / s_idx: [2 * (warps + 1)] scratch; callers alternate the half they pass, so no trailing barrier is needed
    double v = EMD_INF;
    int j = 0x7fffffff;
    for (int t = threadIdx.x; t < M; t += EMD_THREADS) {
        const double k = key[t];
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        s_val[warp] = v;
        s_idx[warp] = j;
    }
    __syncthreads();
    // every thread combines the few warp results itself (same order everywhere -> same answer)
    best = s_val[0];
    best_j = s_idx[0];
#pragma unroll
    for (int w = 1; w < EMD_THREADS / 32; ++w) {
        const double ov = s_val[w];
        const int oj = s_idx[w];
        if (ov < best || (ov == best && oj < best_j)) {
            best = ov;
            best_j = oj;
        }
    }
}

__global__ void __launch_bounds__(EMD_THREADS) emd_kernel(const float* __restrict__ cost, const uint8_t* __restrict__ row_fg,
                                                           const uint32_t* __restrict__ pooled, int P, int64_t m_rows,
                                                           int N, int npw, int t_cap, int16_t* __restrict__ flow_ws,
                                                           double* __restrict__ out, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char emd_smem_raw[];
    EmdSmem s = emd_carve(emd_smem_raw, t_cap, N);
    __shared__ double s_val[2][EMD_THREADS / 32];
    __shared__ int s_idx[2][EMD_THREADS / 32];
    __shared__ int s_T, s_M, s_nreached, s_nnew[2];
    const int tid = threadIdx.x;
    const int64_t lp = blockIdx.x;  // e * P + p
    const int64_t e = lp / P;
    const float* C = cost + e * m_rows * N;
    const uint8_t* fg = row_fg + e * m_rows;
    const uint32_t* pw = pooled + lp * npw;
    int16_t* fT = flow_ws + lp * (int64_t)t_cap * N;  // [M][T] flows (sink-major: a sink's sources are contiguous)

    // ---- index lists (ascending order, like boolean indexing in the reference)
    if (tid == 0) {
        int t = 0;
        for (int64_t r = 0; r < m_rows; ++r)
            if (fg[r]) {
                if (t < t_cap) s.rows[t] = (int)r;
                ++t;
            }
        s_T = t;
        int m = 0;
        for (int b = 0; b < N; ++b)
            if ((pw[b >> 5] >> (b & 31)) & 1u) s.cols[m++] = b;
        s_M = m;
    }
    __syncthreads();
    const int T = s_T, M = s_M;
    if (T == 0 || M == 0) {  // empty marginal: defined as zero transport cost
        if (tid == 0) out[lp] = 1.0;
        return;
    }
    if (T > t_cap) {
        if (tid == 0) {
            out[lp] = nan("");
            atomicMax(status, T);  // tells the host the capacity it needs
        }
        return;
    }
    auto c = [&](int i, int j) -> double { return (double)C[(int64_t)s.rows[i] * N + s.cols[j]]; };

    // ---- initial state: zero flow, u = 0, v_j = min_i c_ij (reduced costs stay >= 0)
    for (int64_t k = tid; k < (int64_t)T * M; k += EMD_THREADS) fT[k] = 0;
    for (int i = tid; i < T; i += EMD_THREADS) {
        s.u[i] = 0.0;
        s.supply[i] = M;
    }
    for (int j = tid; j < M; j += EMD_THREADS) {
        double mn = EMD_INF;
        for (int i = 0; i < T; ++i) mn = fmin(mn, c(i, j));
        s.v[j] = mn;
        s.demand[j] = T;
        s.nfeed[j] = 0;
    }
    __syncthreads();

    for (int r = 0; r < T; ++r) {
        while (s.supply[r] > 0) {  // uniform: shared state only changes between barriers
            EMD_COUNT(0);
            EMD_TIC(t_init);
            // ---- Dijkstra from source r over the sinks
            const double ur = s.u[r];
            for (int j = tid; j < M; j += EMD_THREADS) {
                const double d = c(r, j) - ur - s.v[j];
                s.dist[j] = d;
                s.key[j] = d;
                s.pred_src[j] = r;
            }
            for (int i = tid; i < T; i += EMD_THREADS) s.reached[i] = 0;
            __syncthreads();
            if (tid == 0) {
                s.reached[r] = 1;
                s.dsrc[r] = 0.0;
                s.list[0] = r;
                s_nreached = 1;
            }
            __syncthreads();
            double D;
            int jstar;
            if (tid == 0) s_nnew[0] = s_nnew[1] = 0;
            __syncthreads();
            EMD_TOC(8, t_init);
            for (int step = 0;; ++step) {
                const int pp = step & 1;  // ping-pong scratch / counters: one barrier fewer per step
                EMD_COUNT(1);
                EMD_TIC(t_arg);
                block_argmin(s.key, M, D, jstar, s_val[pp], s_idx[pp]);
                EMD_TOC(9, t_arg);
                if (s.demand[jstar] > 0) break;  // demands only change in the augmentation below: uniform
                if (tid == 0) {
                    s.key[jstar] = EMD_INF;  // scanned
                    s_nnew[pp ^ 1] = 0;      // the other counter is idle during this step
                }
                EMD_TIC(t_col);
                // saturated sink: every source feeding it becomes reachable at distance D (tight backward arcs)
                const int nf = s.nfeed[jstar];
                if (nf != EMD_OVERFLOW) {  // the usual case: the feeders are listed in shared memory
                    if (tid < nf) {
                        const int i = s.feeders[jstar * EMD_INLINE + tid];
                        if (!s.reached[i]) {
                            s.reached[i] = 1;
                            s.dsrc[i] = D;
                            s.pred_sink[i] = jstar;
                            s.newlist[atomicAdd(&s_nnew[pp], 1)] = i;
                            s.list[atomicAdd(&s_nreached, 1)] = i;
                        }
                    }
                } else {
                    const int16_t* frow = fT + (int64_t)jstar * T;
                    for (int i = tid; i < T; i += EMD_THREADS) {
                        if (!s.reached[i] && frow[i] > 0) {
                            s.reached[i] = 1;
                            s.dsrc[i] = D;
                            s.pred_sink[i] = jstar;
                            s.newlist[atomicAdd(&s_nnew[pp], 1)] = i;
                            s.list[atomicAdd(&s_nreached, 1)] = i;
                        }
                    }
                }
                __syncthreads();
                EMD_TOC(10, t_col);
                EMD_TIC(t_rel);
                const int nnew = s_nnew[pp];
#ifdef MARSB200_EMD_PROFILE
                if (tid == 0) atomicAdd((unsigned long long*)&g_emd_prof[2], (unsigned long long)nnew);
#endif
                // relax every unscanned sink against the newly reached sources (usually 0-2 per step)
                for (int j = tid; j < M; j += EMD_THREADS) {
                    if (s.key[j] >= EMD_INF) continue;
                    const float* ccol = C + s.cols[j];
                    double best = EMD_INF;
                    int best_i = -1;
                    int k = 0;
                    for (; k + 4 <= nnew; k += 4) {  // four independent gathers in flight
                        const int i0 = s.newlist[k], i1 = s.newlist[k + 1], i2 = s.newlist[k + 2], i3 = s.newlist[k + 3];
                        const float c0 = ccol[(int64_t)s.rows[i0] * N], c1 = ccol[(int64_t)s.rows[i1] * N];
                        const float c2 = ccol[(int64_t)s.rows[i2] * N], c3 = ccol[(int64_t)s.rows[i3] * N];
                        const double d0 = (double)c0 - s.u[i0], d1 = (double)c1 - s.u[i1];
                        const double d2 = (double)c2 - s.u[i2], d3 = (double)c3 - s.u[i3];
                        if (d0 < best) { best = d0; best_i = i0; }
                        if (d1 < best) { best = d1; best_i = i1; }
                        if (d2 < best) { best = d2; best_i = i2; }
                        if (d3 < best) { best = d3; best_i = i3; }
                    }
                    for (; k < nnew; ++k) {
                        const int i = s.newlist[k];
                        const double d = (double)ccol[(int64_t)s.rows[i] * N] - s.u[i];
                        if (d < best) { best = d; best_i = i; }
                    }
                    if (best_i >= 0) {
                        const double nd = (D + best) - s.v[j];
                        if (nd < s.dist[j]) {
                            s.dist[j] = nd;
                            s.key[j] = nd;
                            s.pred_src[j] = best_i;
                        }
                    }
                }
                __syncthreads();
                EMD_TOC(11, t_rel);
            }
            EMD_TIC(t_dual);
            if (tid == 0) s.key[jstar] = EMD_INF;  // the terminal sink counts as scanned for the dual update
            __syncthreads();
            // ---- dual update: keeps every flow arc tight and all reduced costs non-negative
            const int nreached = s_nreached;
            for (int k = tid; k < nreached; k += EMD_THREADS) {
                const int i = s.list[k];
                s.u[i] += D - s.dsrc[i];
            }
            for (int j = tid; j < M; j += EMD_THREADS)
                if (s.key[j] >= EMD_INF) s.v[j] -= D - s.dist[j];
            __syncthreads();
            EMD_TOC(12, t_dual);
            EMD_TIC(t_aug);
            // ---- augment along the alternating path jstar <- pred_src <- pred_sink <- ... <- r
            if (tid == 0) {
                int delta = min(s.supply[r], s.demand[jstar]);
                int j = jstar;
                while (true) {
                    const int i = s.pred_src[j];
                    if (i == r) break;
                    const int jp = s.pred_sink[i];
                    delta = min(delta, (int)fT[(int64_t)jp * T + i]);
                    j = jp;
                }
                j = jstar;
                while (true) {
                    const int i = s.pred_src[j];
                    const int16_t before = fT[(int64_t)j * T + i];
                    fT[(int64_t)j * T + i] = (int16_t)(before + delta);
                    if (before == 0) {  // new feeder of sink j
                        const int nf = s.nfeed[j];
                        if (nf < EMD_INLINE) {
                            s.feeders[j * EMD_INLINE + nf] = (short)i;
                            s.nfeed[j] = (unsigned char)(nf + 1);
                        } else {
                            s.nfeed[j] = EMD_OVERFLOW;
                        }
                    }
                    if (i == r) break;
                    const int jp = s.pred_sink[i];
                    const int16_t left = (int16_t)(fT[(int64_t)jp * T + i] - delta);
                    fT[(int64_t)jp * T + i] = left;
                    if (left == 0 && s.nfeed[jp] != EMD_OVERFLOW) {  // i no longer feeds sink jp: swap-remove
                        const int nf = s.nfeed[jp];
                        for (int q = 0; q < nf; ++q)
                            if (s.feeders[jp * EMD_INLINE + q] == (short)i) {
                                s.feeders[jp * EMD_INLINE + q] = s.feeders[jp * EMD_INLINE + nf - 1];
                                s.nfeed[jp] = (unsigned char)(nf - 1);
                                break;
                            }
                    }
                    j = jp;
                }
                s.supply[r] -= delta;
                s.demand[jstar] -= delta;
            }
            __syncthreads();
            EMD_TOC(13, t_aug);
        }
    }

    // ---- objective: sum f_ij c_ij / (T M), float64
    double acc = 0.0;
    for (int64_t k = tid; k < (int64_t)T * M; k += EMD_THREADS) {
        const int f = fT[k];
        if (f) {
            const int j = (int)(k / T), i = (int)(k - (int64_t)j * T);
            acc += (double)f * c(i, j);
        }
    }
    acc = warp_sum(acc);
    __syncthreads();
    if ((tid & 31) == 0) s_val[0][tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < EMD_THREADS / 32; ++w) total += s_val[0][w];
        out[lp] = 1.0 - total / ((double)T * (double)M);  // the reference's emd_score = 1 - emd
    }
}

}  // namespace marsb200

using namespace marsb200;

#ifdef MARSB200_EMD_PROFILE
extern "C" int marsb200_debug_emd_profile(long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_emd_prof, sizeof(long long) * 16);
    if (reset) {
        long long z[16] = {0};
        cudaMemcpyToSymbol(g_emd_prof, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" {

int64_t marsb200_emd_workspace_bytes(int E, int P, int N, int t_cap) {
    if (E <= 0 || P <= 0 || N <= 0 || t_cap <= 0) return 0;
    return (int64_t)E * P * t_cap * N * (int64_t)sizeof(int16_t);
}

int marsb200_emd_scores(const float* cost, const uint8_t* row_fg, const uint32_t* pooled, int E, int P, int64_t m_rows,
                        int N, int t_cap, void* workspace, int64_t workspace_bytes, double* out, int32_t* status,
                        void* stream) {
    MARS_REQUIRE(cost && row_fg && pooled && workspace && out && status, "null pointer");
    MARS_REQUIRE(E > 0 && P > 0 && m_rows > 0 && N > 0 && t_cap > 0 && t_cap <= 32767 && N <= 32767, "shape");
    MARS_REQUIRE((int64_t)E * P < (1ll << 31), "too many problems");
    MARS_REQUIRE(workspace_bytes >= marsb200_emd_workspace_bytes(E, P, N, t_cap), "workspace too small");
    const size_t smem = emd_smem_bytes(t_cap, N);
    MARS_REQUIRE(smem <= 200 * 1024, "t_cap + N too large for the shared-memory state");
    cudaStream_t s = as_stream(stream);
    MARS_CUDA_OK(cudaFuncSetAttribute(emd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MARS_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    const int npw = ceil_div(N, 32);
    emd_kernel<<<(unsigned)((int64_t)E * P), EMD_THREADS, smem, s>>>(cost, row_fg, pooled, P, m_rows, N, npw, t_cap,
                                                                    (int16_t*)workspace, out, status);
    MARS_LAUNCH_OK();
    return MARSB200_OK;
}

}  // extern "C"
